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
    const int N = argc > 1 ? atoi(argv[1]) : 8, B = argc > 2 ? atoi(argv[2]) : 256, K = 256, iters = 3;
    const size_t NK = (size_t)N * K;
    std::vector<float> G(NK * NK + NK), P((size_t)B * NK);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX - 0.5f; };
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
    std::vector<int32_t> out(idx.size());
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    long sum = 0;
    for (auto v : out) sum += v;
    printf("checksum %ld\n", sum);
    return (rc || e != cudaSuccess) ? 1 : 0;
}
