// san_encode.cu -- stand-alone driver of the whole C ABI path (prepare, encode, refine, decode, decode backward, classifier
// losses) for compute-sanitizer; random parameters and frames, shapes from the command line:
//   nvcc -o tools/san_encode tools/san_encode.cu -Lquantization_b200 -lmcq -Xlinker -rpath='$ORIGIN/../quantization_b200'
//   compute-sanitizer --tool memcheck tools/san_encode <N> <K> <D> <B> [x_dtype 0|1|2]
// Memory safety and error-free completion are the point; results are only checksummed.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../include/mcq.h"

#define CK(x)                                                                        \
    do {                                                                             \
        int rc__ = (x);                                                              \
        if (rc__) {                                                                  \
            printf("FAILED %s -> %d (%s)\n", #x, rc__, mcq_last_error());           \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8, K = argc > 2 ? atoi(argv[2]) : 256, D = argc > 3 ? atoi(argv[3]) : 96;
    const long B = argc > 4 ? atol(argv[4]) : 300;
    const int xdt = argc > 5 ? atoi(argv[5]) : 0;
    const size_t NK = (size_t)N * K;
    srand(7);
    auto rnd = []() { return (float)rand() / RAND_MAX - 0.5f; };
    std::vector<float> centers(NK * D), w(NK * D), bias(NK), x((size_t)B * D);
    for (auto &v : centers) v = rnd();
    for (size_t i = 0; i < w.size(); ++i) w[i] = 2.0f * centers[i];
    for (auto &v : bias) v = 0.1f * rnd();
    for (auto &v : x) v = 2.0f * rnd();
    float cs = 0.01f, ls = -0.02f;
    float *d_c, *d_w, *d_b, *d_cs, *d_ls;
    cudaMalloc(&d_c, centers.size() * 4); cudaMalloc(&d_w, w.size() * 4); cudaMalloc(&d_b, bias.size() * 4);
    cudaMalloc(&d_cs, 4); cudaMalloc(&d_ls, 4);
    cudaMemcpy(d_c, centers.data(), centers.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_cs, &cs, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ls, &ls, 4, cudaMemcpyHostToDevice);
    const size_t pb = mcq_prepared_bytes(N, K, D), wb = mcq_workspace_bytes(B, D, N, K);
    if (!pb || !wb) { printf("shape rejected: %s\n", mcq_last_error()); return 2; }
    void *blob, *ws;
    cudaMalloc(&blob, pb); cudaMalloc(&ws, wb);
    CK(mcq_prepare(d_c, d_cs, d_w, d_b, d_ls, 10.0f, N, K, D, blob, pb, nullptr));
    void *d_x;
    if (xdt == 0) {
        cudaMalloc(&d_x, x.size() * 4);
        cudaMemcpy(d_x, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    } else {
        std::vector<__half> xh(x.size());
        for (size_t i = 0; i < x.size(); ++i) xh[i] = __float2half(x[i]);
        cudaMalloc(&d_x, x.size() * 2);
        cudaMemcpy(d_x, xh.data(), x.size() * 2, cudaMemcpyHostToDevice);
    }
    const int cols = mcq_packed_cols(N, K);
    unsigned char *d_codes; int64_t *d_idx, *d_idx2;
    cudaMalloc(&d_codes, (size_t)B * cols); cudaMalloc(&d_idx, (size_t)B * N * 8); cudaMalloc(&d_idx2, (size_t)B * N * 8);
    CK(mcq_encode(d_x, xdt == 0 ? MCQ_F32 : MCQ_F16, B, D, N, K, blob, 3, d_codes, MCQ_U8, ws, wb, nullptr));
    CK(mcq_encode(d_x, xdt == 0 ? MCQ_F32 : MCQ_F16, B, D, N, K, blob, 0, d_idx, MCQ_I64, ws, wb, nullptr));
    CK(mcq_refine(d_x, xdt == 0 ? MCQ_F32 : MCQ_F16, B, D, N, K, blob, 2, d_idx, d_idx2, ws, wb, nullptr));
    float *d_out, *d_grad;
    cudaMalloc(&d_out, (size_t)B * D * 4); cudaMalloc(&d_grad, NK * D * 4);
    CK(mcq_decode(d_codes, MCQ_U8, B, cols, N, K, D, blob, d_out, MCQ_F32, nullptr));
    cudaMemset(d_grad, 0, NK * D * 4);
    CK(mcq_decode_backward(d_out, d_idx2, B, N, K, D, d_grad, nullptr));
    if (K <= 256) {
        const long Bp = (B + 127) / 128 * 128;
        float *d_xw, *d_lp, *d_ps, *d_gl, *d_part, *d_g1;
        cudaMalloc(&d_xw, (size_t)Bp * NK * 4); cudaMalloc(&d_lp, 4); cudaMalloc(&d_ps, NK * 4);
        cudaMalloc(&d_gl, (size_t)B * NK * 4); cudaMalloc(&d_part, (size_t)mcq_class_loss_partials() * 4);
        cudaMalloc(&d_g1, 4);
        float one = 1.0f;
        cudaMemcpy(d_g1, &one, 4, cudaMemcpyHostToDevice);
        CK(mcq_class_loss_forward(d_x, xdt == 0 ? MCQ_F32 : MCQ_F16, B, D, N, K, blob, d_idx2, d_xw, d_lp, d_ps, ws, wb,
                                  nullptr));
        CK(mcq_class_loss_backward(d_xw, B, D, N, K, blob, d_idx2, d_g1, d_ps, d_gl, d_part, nullptr));
        float lp = 0;
        cudaMemcpy(&lp, d_lp, 4, cudaMemcpyDeviceToHost);
        printf("logprob_sum %.6f\n", lp);
    }
    if (D <= 1024) {  // fused reconstruction loss, forward and backward; weight-gradient and general products
        float *d_sums, *d_rp, *d_coef, *d_mean;
        cudaMalloc(&d_sums, 8); cudaMalloc(&d_rp, (size_t)mcq_recon_loss_partials() * 4); cudaMalloc(&d_coef, 4);
        cudaMalloc(&d_mean, D * 4);
        cudaMemset(d_mean, 0, D * 4);
        float two = 2.0f;
        cudaMemcpy(d_coef, &two, 4, cudaMemcpyHostToDevice);
        const float *cs = mcq_prepared_scaled_centers(blob, N, K, D);
        CK(mcq_recon_loss_forward(d_x, xdt == 0 ? MCQ_F32 : MCQ_F16, d_idx2, B, N, K, D, cs, d_mean, d_sums, d_rp, nullptr));
        cudaMemset(d_grad, 0, NK * D * 4);
        CK(mcq_recon_loss_backward(d_x, xdt == 0 ? MCQ_F32 : MCQ_F16, d_idx2, B, N, K, D, cs, d_coef, d_grad, nullptr));
        float sums[2] = {0, 0};
        cudaMemcpy(sums, d_sums, 8, cudaMemcpyDeviceToHost);
        printf("recon sums %.4f %.4f\n", sums[0], sums[1]);
        // out (D, D) = d_out^T . d_out over the B frames, and out2 (B, 64k) = d_out . cs[:n]^T
        const size_t wtn = mcq_gemm_tn_workspace_bytes(B, D, D);
        void *d_wtn; float *d_tn;
        cudaMalloc(&d_wtn, wtn); cudaMalloc(&d_tn, (size_t)D * D * 4);
        CK(mcq_gemm_tn(d_out, D, d_out, MCQ_F32, D, B, D, D, d_tn, d_wtn, wtn, nullptr));
        const int nn = (int)(NK / 64 * 64) > 256 ? 256 : (int)(NK / 64 * 64);
        if (nn >= 64) {
            const size_t wnt = mcq_gemm_nt_workspace_bytes(B, nn, D);
            void *d_wnt; float *d_nt;
            cudaMalloc(&d_wnt, wnt); cudaMalloc(&d_nt, (size_t)B * nn * 4);
            CK(mcq_gemm_nt(d_out, D, cs, D, B, nn, D, d_nt, nn, 0, d_wnt, wnt, nullptr));
            CK(mcq_gemm_nt(d_out, D, cs, D, B, nn, D, d_nt, nn, 1, d_wnt, wnt, nullptr));
        }
    }
    {  // host-buffer encode, re-entrant form: pinned host frames / codes, caller-owned staging, a non-default stream
        const int hdt = xdt == 0 ? MCQ_F32 : MCQ_F16;
        const size_t hb = mcq_encode_host_ws_bytes(B, D, N, K, hdt, MCQ_U8);
        void *staging, *hx, *hc;
        cudaStream_t hs;
        cudaMalloc(&staging, hb);
        cudaMallocHost(&hx, x.size() * (xdt == 0 ? 4 : 2));
        cudaMallocHost(&hc, (size_t)B * cols);
        cudaMemcpy(hx, d_x, x.size() * (xdt == 0 ? 4 : 2), cudaMemcpyDeviceToHost);
        cudaStreamCreate(&hs);
        CK(mcq_encode_host_ws(hx, hdt, B, D, N, K, blob, 3, hc, MCQ_U8, staging, hb, hs));
        CK(mcq_encode_host_ws(hx, hdt, B, D, N, K, blob, 3, hc, MCQ_U8, staging, hb / 2 + 4096, hs));  // smaller chunks
        cudaStreamSynchronize(hs);
        std::vector<unsigned char> dc((size_t)B * cols);
        cudaMemcpy(dc.data(), d_codes, dc.size(), cudaMemcpyDeviceToHost);
        long diff = 0;
        for (size_t i = 0; i < dc.size(); ++i) diff += dc[i] != ((unsigned char *)hc)[i];
        printf("encode_host_ws vs encode: %ld differing bytes\n", diff);
        if (diff) return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned char> codes((size_t)B * cols);
    cudaMemcpy(codes.data(), d_codes, codes.size(), cudaMemcpyDeviceToHost);
    long sum = 0;
    for (auto v : codes) sum += v;
    printf("N=%d K=%d D=%d B=%ld dtype=%d: cuda=%s codes checksum %ld\n", N, K, D, B, xdt, cudaGetErrorString(e), sum);
    return e == cudaSuccess ? 0 : 1;
}
