// san_jcl.cu -- stand-alone driver of the JointCodebookLoss stages (mcq_jcl_*) for compute-sanitizer:
//   nvcc -o tools/san_jcl tools/san_jcl.cu -Lquantization_b200 -lmcq -Xlinker -rpath=$PWD/quantization_b200
//   compute-sanitizer --tool memcheck tools/san_jcl 8 256 512 1000 0     (N K H B codes_dtype[0=u8,1=i64,2=i32])
// Random inputs (negative = padded codes for the signed types); the point is memory safety, not the result.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../include/mcq.h"

#define CK(call)                                                             \
    do {                                                                     \
        int rc__ = (call);                                                   \
        if (rc__ != 0) {                                                     \
            printf("%s failed: %d (%s)\n", #call, rc__, mcq_last_error());   \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8, K = argc > 2 ? atoi(argv[2]) : 256, H = argc > 3 ? atoi(argv[3]) : 512;
    const int B = argc > 4 ? atoi(argv[4]) : 1000, dt = argc > 5 ? atoi(argv[5]) : 0;
    srand(1);
    std::vector<float> hid((size_t)B * H), emb((size_t)(N - 1) * K * H), gact((size_t)N * B * H), logits((size_t)B * N * K),
        bias((size_t)N * K);
    for (auto &v : hid) v = rand() / (float)RAND_MAX - 0.5f;
    for (auto &v : emb) v = rand() / (float)RAND_MAX - 0.5f;
    for (auto &v : gact) v = rand() / (float)RAND_MAX - 0.5f;
    for (auto &v : logits) v = 4.0f * (rand() / (float)RAND_MAX - 0.5f);
    for (auto &v : bias) v = rand() / (float)RAND_MAX - 0.5f;
    const size_t esz = dt == 0 ? 1 : (dt == 1 ? 8 : 4);
    std::vector<unsigned char> codes((size_t)B * N * esz);
    for (size_t i = 0; i < (size_t)B * N; ++i) {
        long long c = rand() % K;
        if (dt != 0 && (i / N) % 5 == 0) c = -100;
        if (dt == 0) codes[i] = (unsigned char)c;
        else if (dt == 1) ((long long *)codes.data())[i] = c;
        else ((int *)codes.data())[i] = (int)c;
    }
    float *d_hid, *d_emb, *d_act, *d_gact, *d_gh, *d_gemb, *d_logits, *d_bias, *d_row, *d_sums, *d_part;
    void *d_codes;
    cudaMalloc(&d_hid, hid.size() * 4); cudaMalloc(&d_emb, emb.size() * 4); cudaMalloc(&d_act, gact.size() * 4);
    cudaMalloc(&d_gact, gact.size() * 4); cudaMalloc(&d_gh, hid.size() * 4); cudaMalloc(&d_gemb, emb.size() * 4);
    cudaMalloc(&d_logits, logits.size() * 4); cudaMalloc(&d_bias, bias.size() * 4); cudaMalloc(&d_row, (size_t)B * N * 4);
    cudaMalloc(&d_sums, 8); cudaMalloc(&d_part, (size_t)mcq_jcl_partials() * 4); cudaMalloc(&d_codes, codes.size());
    cudaMemcpy(d_hid, hid.data(), hid.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_emb, emb.data(), emb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_gact, gact.data(), gact.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_logits, logits.data(), logits.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_codes, codes.data(), codes.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_gemb, 0, emb.size() * 4);
    const int cdt = dt == 0 ? MCQ_U8 : (dt == 1 ? MCQ_I64 : MCQ_I32);
    CK(mcq_jcl_hidden_forward(d_hid, d_codes, cdt, B, N, K, H, d_emb, 0.7f, d_act, nullptr));
    CK(mcq_jcl_hidden_backward(d_gact, d_act, d_codes, cdt, B, N, K, H, 0.7f, d_gh, d_gemb, nullptr));
    CK(mcq_jcl_cross_entropy(d_logits, d_bias, d_codes, cdt, B, N, K, -100, 1, d_row, d_sums, d_part, nullptr));
    float sums[2];
    cudaError_t e = cudaMemcpy(sums, d_sums, 8, cudaMemcpyDeviceToHost);
    std::vector<float> gh(hid.size());
    cudaMemcpy(gh.data(), d_gh, gh.size() * 4, cudaMemcpyDeviceToHost);
    double cs = 0;
    for (float v : gh) cs += v;
    printf("loss %.4f rows %.0f grad_hidden checksum %.4f (%s)\n", sums[0], sums[1], cs, cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
