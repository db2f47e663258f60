// Host harness for tests/test_car_env_host_cpu.py: the car env's per-environment step / reset logic
// (mobrob_b200/csrc/env_car.cuh on car_dyn.cuh, the functions car_step_kernel / car_reset_kernel call) compiled for the
// HOST and driven like mr_env_reset + T x mr_env_step.  Same file format as point_env_host.cu with 26-float rows.
#include <vector>

#include "../../mobrob_b200/csrc/env_car.cuh"

namespace mr {
void set_error(const char*, ...) {}
void count_launch(uint64_t) {}
}  // namespace mr

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 3;
    int64_t hdr[4];
    if (fread(hdr, 8, 4, fi) != 4) return 4;
    const int64_t n = hdr[0], T = hdr[1];
    constexpr int O = mr::car::OBS;
    std::vector<uint64_t> pcg_init((size_t)n * 4), pcg_goal((size_t)n * 4);
    std::vector<int64_t> engine_seed((size_t)n);
    std::vector<float> act((size_t)T * n * 2);
    if (fread(pcg_init.data(), 8, pcg_init.size(), fi) != pcg_init.size()) return 4;
    if (fread(pcg_goal.data(), 8, pcg_goal.size(), fi) != pcg_goal.size()) return 4;
    if (fread(engine_seed.data(), 8, engine_seed.size(), fi) != engine_seed.size()) return 4;
    if (fread(act.data(), 4, act.size(), fi) != act.size()) return 4;
    std::vector<float2> body_xy((size_t)n);
    std::vector<double> psi0((size_t)n);
    std::vector<int32_t> counts((size_t)n * 2, 0);
    const float spaces[8] = {-1.f, -1.f, 1.f, 1.f, -2.f, -2.f, 2.f, 2.f};   // wrapper.py:250-264
    mr::EnvCold cold{pcg_init.data(), pcg_goal.data(), engine_seed.data(), body_xy.data(), psi0.data(), counts.data(), spaces};
    mr::EnvCfg cfg{};
    cfg.time_limit = (int)hdr[2];
    cfg.terminate_on_goal = (int)hdr[3];
    cfg.pk = mr::point::make_k();
    cfg.obs_flags = 0;
    const mr::car::Consts K = mr::car::make_consts();
    std::vector<double> scratch(mr::car::SCRATCH_DOUBLES);
    mr::car::Scratch S{};
    S.host = scratch.data();
    std::vector<mr::CarHot> hot((size_t)n);
    std::vector<float> obs((size_t)n * O), tobs((size_t)n * O, 0.f), rew((size_t)n);
    std::vector<uint8_t> done((size_t)n), trunc((size_t)n);
    std::vector<double> ep_r((size_t)n, 0.0);
    std::vector<int32_t> ep_l((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {   // mr_env_reset(first = 1): car_reset_kernel
        mr::CarHot h{};
        mr::car_reset(h, cold, i, true);
        mr::car::sensors(K, h.s, (double)h.cx, (double)h.cz, h.gx, h.gy, true, &obs[(size_t)i * O], S);
        hot[i] = h;
    }
    fwrite(obs.data(), 4, obs.size(), fo);
    for (int64_t t = 0; t < T; ++t) {
        for (int64_t i = 0; i < n; ++i) {   // car_step_kernel, one thread per env
            const float* a = &act[((size_t)t * n + i) * 2];
            float o[O], to[O];
            const mr::StepResult r = mr::car_env_step(K, hot[i], cold, i, a[0], a[1], cfg, true, o, to, S);
            for (int k = 0; k < O; ++k) obs[(size_t)i * O + k] = o[k];
            rew[i] = r.rew; done[i] = r.done; trunc[i] = r.trunc;
            if (r.done) {
                for (int k = 0; k < O; ++k) tobs[(size_t)i * O + k] = to[k];
                ep_r[i] = r.ep_r; ep_l[i] = r.ep_l;
            }
        }
        fwrite(obs.data(), 4, obs.size(), fo);
        fwrite(rew.data(), 4, rew.size(), fo);
        fwrite(done.data(), 1, done.size(), fo);
        fwrite(trunc.data(), 1, trunc.size(), fo);
        fwrite(tobs.data(), 4, tobs.size(), fo);
        fwrite(ep_r.data(), 8, ep_r.size(), fo);
        fwrite(ep_l.data(), 4, ep_l.size(), fo);
    }
    fwrite(counts.data(), 4, counts.size(), fo);
    fclose(fo);
    return 0;
}
