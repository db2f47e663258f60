// Host harness for tests/test_point_env_host_cpu.py: the per-environment step / reset logic the env-step and rollout
// kernels call (mobrob_b200/csrc/env_point.cuh: EnvWrapper.step / reward / reached / reset, TimeLimit, Monitor,
// auto-reset, the PCG64 / MT19937 restatements of the reference's random streams) compiled for the HOST and driven
// like mr_env_reset + T x mr_env_step.
//   in : int64 n, T, time_limit, terminate_on_goal | uint64 pcg_init[n][4] | uint64 pcg_goal[n][4] | int64 engine_seed[n]
//        | float act[T][n][2]
//   out: float obs0[n][14] | per step: float obs[n][14], float rew[n], uint8 done[n], uint8 trunc[n],
//        float term_obs[n][14] (rows of finished envs), double ep_r[n], int32 ep_l[n] | int32 counts[n][2]
#include <vector>

#include "../../mobrob_b200/csrc/env_point.cuh"

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
    std::vector<mr::PointHot> hot((size_t)n);
    std::vector<float> obs((size_t)n * 14), tobs((size_t)n * 14, 0.f), rew((size_t)n);
    std::vector<uint8_t> done((size_t)n), trunc((size_t)n);
    std::vector<double> ep_r((size_t)n, 0.0);
    std::vector<int32_t> ep_l((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {   // mr_env_reset(first = 1): point_reset_kernel
        mr::PointHot h{};
        mr::point_reset(h, cold, i, true);
        mr::point::sensors(h.d, (double)h.cx, (double)h.cz, h.gx, h.gy, &obs[(size_t)i * 14]);
        hot[i] = h;
    }
    fwrite(obs.data(), 4, obs.size(), fo);
    for (int64_t t = 0; t < T; ++t) {
        for (int64_t i = 0; i < n; ++i) {   // point_step_kernel, one thread per env
            const float* a = &act[((size_t)t * n + i) * 2];
            float o[14], to[14];
            const mr::StepResult r = mr::point_env_step(hot[i], cold, i, a[0], a[1], cfg, o, to);
            for (int k = 0; k < 14; ++k) obs[(size_t)i * 14 + k] = o[k];
            rew[i] = r.rew; done[i] = r.done; trunc[i] = r.trunc;
            if (r.done) {
                for (int k = 0; k < 14; ++k) tobs[(size_t)i * 14 + k] = to[k];
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
