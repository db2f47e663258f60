// Host harness for tests/test_point_dyn_host_cpu.py: runs the SAME integrator / sensor functions the env-step
// and rollout kernels call (mobrob_b200/csrc/point_dyn.cuh, __host__ __device__) on the CPU, so that the
// reformulated physics (linear form of the angular acceleration, loop over the heading increment, Taylor
// rotation, constant-bank sincos, cold large-angle path) is checked against the oracle without a GPU.
//   in : int64 n, int64 T | double state[n][6] (px py psi vx vy om) | float goal[n][2] | float act[T][n][2]
//   out: double state[n][6] after T env steps | float obs[T][n][14]
#include <vector>

#include "../../mobrob_b200/csrc/point_dyn.cuh"

// common.cuh declares these; the library defines them in env.cu
namespace mr {
void set_error(const char*, ...) {}
void count_launch(uint64_t) {}
}  // namespace mr

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 3;
    int64_t n = 0, T = 0;
    if (fread(&n, 8, 1, fi) != 1 || fread(&T, 8, 1, fi) != 1) return 4;
    std::vector<double> st((size_t)n * 6);
    std::vector<float> goal((size_t)n * 2), act((size_t)T * n * 2), obs((size_t)T * n * 14);
    if (fread(st.data(), 8, st.size(), fi) != st.size()) return 4;
    if (fread(goal.data(), 4, goal.size(), fi) != goal.size()) return 4;
    if (fread(act.data(), 4, act.size(), fi) != act.size()) return 4;
    const mr::point::K k = mr::point::make_k();
    for (int64_t i = 0; i < n; ++i) {
        mr::point::Dyn d{st[6 * i], st[6 * i + 1], st[6 * i + 2], st[6 * i + 3], st[6 * i + 4], st[6 * i + 5]};
        for (int64_t t = 0; t < T; ++t) {
            const float* a = &act[((size_t)t * n + i) * 2];
            const float cx = a[0] < -1.f ? -1.f : (a[0] > 1.f ? 1.f : a[0]);   // engine.py:1401-1405
            const float cz = a[1] < -1.f ? -1.f : (a[1] > 1.f ? 1.f : a[1]);
            double c, s;
            mr::point::substeps(k, d, (double)cx, (double)cz, c, s);
            const double dx = (double)goal[2 * i] - d.px, dy = (double)goal[2 * i + 1] - d.py;
            mr::point::sensors_cs(k, d, c, s, (double)cx, (double)cz, goal[2 * i], goal[2 * i + 1],
                                  &obs[((size_t)t * n + i) * 14], sqrt(dx * dx + dy * dy));
        }
        st[6 * i] = d.px; st[6 * i + 1] = d.py; st[6 * i + 2] = d.psi;
        st[6 * i + 3] = d.vx; st[6 * i + 4] = d.vy; st[6 * i + 5] = d.om;
    }
    fwrite(st.data(), 8, st.size(), fo);
    fwrite(obs.data(), 4, obs.size(), fo);
    fclose(fo);
    return 0;
}
