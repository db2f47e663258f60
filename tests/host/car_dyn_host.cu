// Host harness for tests/test_car_dyn_host_cpu.py: runs the SAME substep / sensor routines the car kernels call
// (mobrob_b200/csrc/car_dyn.cuh, __host__ __device__; the contact scratch is a plain array on the host) on the CPU, so
// that the closed-form Delassus matrix, the warm-started projected Gauss-Seidel sweeps and the gyrostat dynamics are
// checked against the oracle's matrix-free restatement without a GPU.
//   in : int64 n, int64 T, int64 contacts | double state[n][24] (p3 quat4 v3 w3 th2 s2 qb4 wb3: car::State order)
//        | float goal[n][2] | float act[T][n][2]
//   out: double state[n][24] after T env steps | float obs[T][n][26]
#include <vector>

#include "../../mobrob_b200/csrc/car_dyn.cuh"

namespace mr {
void set_error(const char*, ...) {}
void count_launch(uint64_t) {}
}  // namespace mr

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 3;
    int64_t n = 0, T = 0, contacts = 1;
    if (fread(&n, 8, 1, fi) != 1 || fread(&T, 8, 1, fi) != 1 || fread(&contacts, 8, 1, fi) != 1) return 4;
    static_assert(sizeof(mr::car::State) == mr::car::NSTATE * sizeof(double), "State is 24 packed doubles");
    std::vector<mr::car::State> st((size_t)n);
    std::vector<float> goal((size_t)n * 2), act((size_t)T * n * 2), obs((size_t)T * n * mr::car::OBS);
    if (fread(st.data(), sizeof(mr::car::State), st.size(), fi) != st.size()) return 4;
    if (fread(goal.data(), 4, goal.size(), fi) != goal.size()) return 4;
    if (fread(act.data(), 4, act.size(), fi) != act.size()) return 4;
    const mr::car::Consts K = mr::car::make_consts();
    std::vector<double> scratch(mr::car::SCRATCH_DOUBLES);
    mr::car::Scratch S{};
    S.host = scratch.data();
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t t = 0; t < T; ++t) {
            const float* a = &act[((size_t)t * n + i) * 2];
            const float cx = a[0] < -1.f ? -1.f : (a[0] > 1.f ? 1.f : a[0]);   // engine.py:1401-1405
            const float cz = a[1] < -1.f ? -1.f : (a[1] > 1.f ? 1.f : a[1]);
            for (int k = 0; k < mr::car::FRAME_SKIP; ++k)   // car_env_step (env_car.cuh): the first substep is cold
                mr::car::substep(K, st[i], (double)cx, (double)cz, contacts != 0, S, k > 0);
            mr::car::sensors(K, st[i], (double)cx, (double)cz, goal[2 * i], goal[2 * i + 1], contacts != 0,
                             &obs[((size_t)t * n + i) * mr::car::OBS], S);
        }
    }
    fwrite(st.data(), sizeof(mr::car::State), st.size(), fo);
    fwrite(obs.data(), 4, obs.size(), fo);
    fclose(fo);
    return 0;
}
