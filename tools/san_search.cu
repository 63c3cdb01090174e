// san_search.cu -- tiny stand-alone driver of mcq_search for compute-sanitizer (memcheck / racecheck / synccheck):
//   nvcc -o /tmp/san_search tools/san_search.cu -Lquantization_b200 -lmcq -Xlinker -rpath=$PWD/quantization_b200
//   compute-sanitizer --tool memcheck /tmp/san_search 4 512
// Random symmetric "Gram" table and random P: the point is memory safety of the search kernels, not the result.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../include/mcq.h"

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8, B = argc > 2 ? atoi(argv[2]) : 256, iters = 3;
    const int K = argc > 4 ? atoi(argv[4]) : 256;
    const int mode = argc > 3 ? atoi(argv[3]) : 0;  // 1: heavily quantised values (many exactly equal scores), 2: all zero
    const size_t NK = (size_t)N * K;
    std::vector<float> G(NK * NK + NK), P((size_t)B * NK);
    srand(1);
    auto rnd = [mode]() {
        if (mode == 2) return 0.0f;
        if (mode == 1) return (float)(rand() % 5 - 2) * 0.25f;
        return (float)rand() / RAND_MAX - 0.5f;
    };
    for (size_t r = 0; r < NK; ++r)
        for (size_t c = r; c < NK; ++c) {
            float v = r == c ? 1.0f + rnd() : 0.1f * rnd();
            G[r * NK + c] = v;
            G[c * NK + r] = v;
        }
    for (size_t r = 0; r < NK; ++r) G[NK * NK + r] = G[r * NK + r];
    for (auto &v : P) v = rnd();
    std::vector<int32_t> idx((size_t)B * N);
    for (auto &v : idx) v = rand() % K;
    float *dG, *dP;
    int32_t *dI, *dO;
    cudaMalloc(&dG, G.size() * 4);
    cudaMalloc(&dP, P.size() * 4);
    cudaMalloc(&dI, idx.size() * 4);
    cudaMalloc(&dO, idx.size() * 4);
    cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dI, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
    int rc = mcq_search(dP, dG, B, N, K, iters, dI, dO, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mcq_search rc=%d (%s) cuda=%s\n", rc, mcq_last_error(), cudaGetErrorString(e));
    std::vector<int32_t> out(idx.size()), ref(idx.size());
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    // the generic first-version kernel on the same input must agree exactly (ties included)
    setenv("MCQ_SEARCH", "v1", 1);
    int rc1 = mcq_search(dP, dG, B, N, K, iters, dI, dO, nullptr);
    cudaError_t e1 = cudaDeviceSynchronize();
    cudaMemcpy(ref.data(), dO, ref.size() * 4, cudaMemcpyDeviceToHost);
    long sum = 0, bad = 0;
    for (size_t i = 0; i < out.size(); ++i) {
        sum += out[i];
        bad += out[i] != ref[i];
    }
    printf("checksum %ld, entries differing from v1: %ld (rc1=%d %s)\n", sum, bad, rc1, cudaGetErrorString(e1));
    return (rc || rc1 || bad || e != cudaSuccess || e1 != cudaSuccess) ? 1 : 0;
}
