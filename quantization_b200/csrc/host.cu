// host.cu -- mcq_encode_host: Quantizer.encode (quantization.py:244-275) for a caller that holds HOST buffers.
// Frames are streamed through the device in chunks on three streams (H2D, compute, D2H) with double buffering, so
// the PCIe copies of chunk i+1 / i-1 overlap the kernels of chunk i.  Device buffers are cached per device.
#include <mutex>

#include "common.cuh"

namespace mcq {

namespace {

struct HostCtx {
    bool init = false;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    void *d_x[2] = {nullptr, nullptr};
    void *d_codes[2] = {nullptr, nullptr};
    void *ws = nullptr;
    size_t x_cap[2] = {0, 0}, codes_cap[2] = {0, 0}, ws_cap = 0;
};

constexpr int MAX_DEV = 16;
HostCtx g_ctx[MAX_DEV];
std::mutex g_mu[MAX_DEV];

int grow(void **p, size_t *cap, size_t need) {
    if (*cap >= need) return MCQ_OK;
    if (*p) MCQ_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    MCQ_CUDA(cudaMalloc(p, need));
    *cap = need;
    return MCQ_OK;
}

// The body of mcq_encode_host with the device selected and the context locked; may return early on any error (the
// caller synchronises the streams and restores the device).
int encode_host_locked(HostCtx &c, const void *x_host, int x_dtype, int64_t B, int D, int N, int K,
                       const void *prepared, int iters, void *codes_host, int codes_dtype) {
    int rc = MCQ_OK;
    if (!c.init) {
        MCQ_CUDA(cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking));
        MCQ_CUDA(cudaStreamCreateWithFlags(&c.s_cmp, cudaStreamNonBlocking));
        MCQ_CUDA(cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            MCQ_CUDA(cudaEventCreateWithFlags(&c.ev_in[k], cudaEventDisableTiming));
            MCQ_CUDA(cudaEventCreateWithFlags(&c.ev_cmp[k], cudaEventDisableTiming));
            MCQ_CUDA(cudaEventCreateWithFlags(&c.ev_out[k], cudaEventDisableTiming));
        }
        c.init = true;
    }
    const size_t xelt = x_dtype == MCQ_F32 ? 4 : 2;
    const int ncols = codes_dtype == MCQ_U8 ? mcq_packed_cols(N, K) : N;
    const size_t celt = codes_dtype == MCQ_U8 ? 1 : (codes_dtype == MCQ_I64 ? 8 : 4);
    int64_t Bc = 148 * 128 * 2;  // 37,888 frames per chunk
    if (Bc > B) Bc = (int64_t)align_up((size_t)B, 128);
    const size_t ws_need = mcq_workspace_bytes(Bc, D, N, K);
    if ((rc = grow(&c.ws, &c.ws_cap, ws_need))) return rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = grow(&c.d_x[k], &c.x_cap[k], (size_t)Bc * D * xelt))) return rc;
        if ((rc = grow(&c.d_codes[k], &c.codes_cap[k], (size_t)Bc * ncols * celt))) return rc;
    }
    // the caller's stream (legacy default) may still be producing `prepared`: order after it
    MCQ_CUDA(cudaStreamSynchronize(nullptr));
    int64_t chunk = 0;
    for (int64_t b0 = 0; b0 < B; b0 += Bc, ++chunk) {
        const int k = (int)(chunk & 1);
        const int64_t nb = B - b0 < Bc ? B - b0 : Bc;
        // d_x[k] was last read by the compute of chunk-2, d_codes[k] by the D2H of chunk-2
        MCQ_CUDA(cudaStreamWaitEvent(c.s_in, c.ev_cmp[k], 0));
        MCQ_CUDA(cudaMemcpyAsync(c.d_x[k], (const char *)x_host + (size_t)b0 * D * xelt, (size_t)nb * D * xelt,
                                 cudaMemcpyHostToDevice, c.s_in));
        MCQ_CUDA(cudaEventRecord(c.ev_in[k], c.s_in));
        MCQ_CUDA(cudaStreamWaitEvent(c.s_cmp, c.ev_in[k], 0));
        MCQ_CUDA(cudaStreamWaitEvent(c.s_cmp, c.ev_out[k], 0));
        if ((rc = mcq_encode(c.d_x[k], x_dtype, nb, D, N, K, prepared, iters, c.d_codes[k], codes_dtype, c.ws, c.ws_cap,
                             c.s_cmp)))
            return rc;
        MCQ_CUDA(cudaEventRecord(c.ev_cmp[k], c.s_cmp));
        MCQ_CUDA(cudaStreamWaitEvent(c.s_out, c.ev_cmp[k], 0));
        MCQ_CUDA(cudaMemcpyAsync((char *)codes_host + (size_t)b0 * ncols * celt, c.d_codes[k],
                                 (size_t)nb * ncols * celt, cudaMemcpyDeviceToHost, c.s_out));
        MCQ_CUDA(cudaEventRecord(c.ev_out[k], c.s_out));
    }
    return MCQ_OK;
}

// Layout of the caller-owned staging buffer of mcq_encode_host_ws for chunks of Bc frames: two frame buffers, two
// code buffers, one mcq_encode workspace (every region 1024-byte aligned).
struct HostWs {
    int64_t Bc;
    size_t off_x[2], off_codes[2], off_ws, ws_bytes, bytes;
};

HostWs host_ws_layout(int64_t Bc, int D, int N, int K, int x_dtype, int codes_dtype) {
    HostWs L;
    L.Bc = Bc;
    const size_t xelt = x_dtype == MCQ_F32 ? 4 : 2;
    const int ncols = codes_dtype == MCQ_U8 ? mcq_packed_cols(N, K) : N;
    const size_t celt = codes_dtype == MCQ_U8 ? 1 : (codes_dtype == MCQ_I64 ? 8 : 4);
    size_t off = 0;
    for (int k = 0; k < 2; ++k) {
        L.off_x[k] = off;
        off += align_up((size_t)Bc * D * xelt, 1024);
    }
    for (int k = 0; k < 2; ++k) {
        L.off_codes[k] = off;
        off += align_up((size_t)Bc * ncols * celt, 1024);
    }
    L.off_ws = off;
    L.ws_bytes = mcq_workspace_bytes(Bc, D, N, K);
    L.bytes = off + align_up(L.ws_bytes, 1024);
    return L;
}

constexpr int64_t HOST_CHUNK = 148 * 128 * 4;  // 75,776 frames: one chunk of mcq_encode (four waves of GEMM tiles)
constexpr int64_t HOST_FIRST = 148 * 128;      // one wave: the unit of the chunk-size ramp (see the loop)

}  // namespace

}  // namespace mcq

using namespace mcq;

extern "C" int mcq_encode_host(const void *x_host, int x_dtype, int64_t B, int D, int N, int K, const void *prepared,
                               int iters, void *codes_host, int codes_dtype, int device) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B < 0 || iters < 0 || x_dtype < 0 || x_dtype > 2 || codes_dtype < 0 || codes_dtype > 2 || device < 0 ||
        device >= MAX_DEV) {
        set_error("mcq_encode_host: bad argument");
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!x_host || !prepared || !codes_host) {
        set_error("mcq_encode_host: null pointer");
        return MCQ_EINVAL;
    }
    std::lock_guard<std::mutex> lock(g_mu[device]);
    int prev_dev = 0;
    MCQ_CUDA(cudaGetDevice(&prev_dev));
    MCQ_CUDA(cudaSetDevice(device));
    // ONE exit path: whatever happens below, nothing is left in flight on the three streams (they touch the caller's
    // host buffers and the cached device buffers, which a later call may free and re-grow) and the caller's current
    // device is restored.
    struct Cleanup {
        HostCtx &c;
        int prev_dev;
        ~Cleanup() {
            if (c.s_in) cudaStreamSynchronize(c.s_in);
            if (c.s_cmp) cudaStreamSynchronize(c.s_cmp);
            if (c.s_out) cudaStreamSynchronize(c.s_out);
            cudaSetDevice(prev_dev);
        }
    } cleanup{g_ctx[device], prev_dev};
    rc = encode_host_locked(g_ctx[device], x_host, x_dtype, B, D, N, K, prepared, iters, codes_host, codes_dtype);
    if (rc == MCQ_OK) {
        // report asynchronous failures of the last copies / kernels instead of swallowing them in the guard
        cudaError_t e = cudaStreamSynchronize(g_ctx[device].s_out);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx[device].s_cmp);
        if (e != cudaSuccess) rc = cuda_fail(e, "mcq_encode_host: stream synchronisation");
    }
    return rc;
}

extern "C" size_t mcq_encode_host_ws_bytes(int64_t num_frames, int D, int N, int K, int x_dtype, int codes_dtype) {
    if (check_shape(N, K, D) || num_frames < 1 || x_dtype < 0 || x_dtype > 2 || codes_dtype < 0 || codes_dtype > 2)
        return 0;
    int64_t Bc = HOST_CHUNK;
    if (Bc > num_frames) Bc = (int64_t)align_up((size_t)num_frames, 128);
    return host_ws_layout(Bc, D, N, K, x_dtype, codes_dtype).bytes;
}

// The re-entrant form: the caller owns the device staging buffer and the stream; nothing is allocated, no state is
// kept between calls and the host is never blocked.  The copies and kernels run on two helper streams created and
// released inside the call (stream / event destruction is deferred by the driver until their work has drained), ordered
// after everything already on `stream`, and `stream` is made to wait for the last of them.
extern "C" int mcq_encode_host_ws(const void *x_host, int x_dtype, int64_t B, int D, int N, int K, const void *prepared,
                                  int iters, void *codes_host, int codes_dtype, void *staging, size_t staging_bytes,
                                  void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B < 0 || iters < 0 || x_dtype < 0 || x_dtype > 2 || codes_dtype < 0 || codes_dtype > 2) {
        set_error("mcq_encode_host_ws: bad argument");
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!x_host || !prepared || !codes_host || !staging) {
        set_error("mcq_encode_host_ws: null pointer");
        return MCQ_EINVAL;
    }
    // the largest chunk (a multiple of 128 frames, at most HOST_CHUNK) whose layout fits the caller's buffer
    int64_t Bc = HOST_CHUNK;
    if (Bc > B) Bc = (int64_t)align_up((size_t)B, 128);
    while (Bc >= 128 && host_ws_layout(Bc, D, N, K, x_dtype, codes_dtype).bytes > staging_bytes) Bc -= 128;
    if (Bc < 128) {
        set_error("mcq_encode_host_ws: staging buffer of %zu bytes is too small (mcq_encode_host_ws_bytes)", staging_bytes);
        return MCQ_EINVAL;
    }
    const HostWs L = host_ws_layout(Bc, D, N, K, x_dtype, codes_dtype);
    const size_t xelt = x_dtype == MCQ_F32 ? 4 : 2;
    const int ncols = codes_dtype == MCQ_U8 ? mcq_packed_cols(N, K) : N;
    const size_t celt = codes_dtype == MCQ_U8 ? 1 : (codes_dtype == MCQ_I64 ? 8 : 4);
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)staging;

    struct Res {  // released on every exit path; destruction of busy streams / events is deferred by the driver
        cudaStream_t s_in = nullptr, s_out = nullptr;
        cudaEvent_t ev_start = nullptr, ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr},
                    ev_out[2] = {nullptr, nullptr};
        ~Res() {
            for (int k = 0; k < 2; ++k) {
                if (ev_in[k]) cudaEventDestroy(ev_in[k]);
                if (ev_cmp[k]) cudaEventDestroy(ev_cmp[k]);
                if (ev_out[k]) cudaEventDestroy(ev_out[k]);
            }
            if (ev_start) cudaEventDestroy(ev_start);
            if (s_in) cudaStreamDestroy(s_in);
            if (s_out) cudaStreamDestroy(s_out);
        }
    } r;
    MCQ_CUDA(cudaStreamCreateWithFlags(&r.s_in, cudaStreamNonBlocking));
    MCQ_CUDA(cudaStreamCreateWithFlags(&r.s_out, cudaStreamNonBlocking));
    MCQ_CUDA(cudaEventCreateWithFlags(&r.ev_start, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        MCQ_CUDA(cudaEventCreateWithFlags(&r.ev_in[k], cudaEventDisableTiming));
        MCQ_CUDA(cudaEventCreateWithFlags(&r.ev_cmp[k], cudaEventDisableTiming));
        MCQ_CUDA(cudaEventCreateWithFlags(&r.ev_out[k], cudaEventDisableTiming));
    }
    // everything already enqueued on the caller's stream (e.g. mcq_prepare, earlier users of `staging`) comes first
    MCQ_CUDA(cudaEventRecord(r.ev_start, st));
    MCQ_CUDA(cudaStreamWaitEvent(r.s_in, r.ev_start, 0));
    MCQ_CUDA(cudaStreamWaitEvent(r.s_out, r.ev_start, 0));
    int64_t chunk = 0;
    for (int64_t b0 = 0, nb = 0; b0 < B; b0 += nb, ++chunk) {
        const int k = (int)(chunk & 1);
        // chunk sizes ramp 1, 2, 4, ... 4, 2, 1 waves: the first copy and the last compute are the only stages nothing
        // overlaps (the last matters when the host link is the bottleneck: 8 ranks on one host)
        const int64_t left = B - b0, w = HOST_FIRST;
        nb = left < Bc ? left : Bc;
        if (chunk < 2 && nb > (w << chunk)) nb = w << chunk;
        if (left > w && left <= 3 * w) {
            if (nb > left - w) nb = left - w;                            // ... then one last wave
        } else if (left > 3 * w && left <= 7 * w) {
            if (nb > left - 3 * w) nb = left - 3 * w;                    // ... then 2 + 1 waves
        }
        nb = (nb + 127) / 128 * 128;  // whole GEMM tiles, except at the very end
        if (nb > left) nb = left;
        if (chunk >= 2) MCQ_CUDA(cudaStreamWaitEvent(r.s_in, r.ev_cmp[k], 0));  // x buffer k was read by chunk - 2
        MCQ_CUDA(cudaMemcpyAsync(base + L.off_x[k], (const char *)x_host + (size_t)b0 * D * xelt, (size_t)nb * D * xelt,
                                 cudaMemcpyHostToDevice, r.s_in));
        MCQ_CUDA(cudaEventRecord(r.ev_in[k], r.s_in));
        MCQ_CUDA(cudaStreamWaitEvent(st, r.ev_in[k], 0));
        if (chunk >= 2) MCQ_CUDA(cudaStreamWaitEvent(st, r.ev_out[k], 0));  // code buffer k was copied out by chunk - 2
        if ((rc = mcq_encode(base + L.off_x[k], x_dtype, nb, D, N, K, prepared, iters, base + L.off_codes[k], codes_dtype,
                             base + L.off_ws, L.ws_bytes, st)))
            return rc;
        MCQ_CUDA(cudaEventRecord(r.ev_cmp[k], st));
        MCQ_CUDA(cudaStreamWaitEvent(r.s_out, r.ev_cmp[k], 0));
        MCQ_CUDA(cudaMemcpyAsync((char *)codes_host + (size_t)b0 * ncols * celt, base + L.off_codes[k],
                                 (size_t)nb * ncols * celt, cudaMemcpyDeviceToHost, r.s_out));
        MCQ_CUDA(cudaEventRecord(r.ev_out[k], r.s_out));
    }
    // the caller's stream completes when the last codes are in host memory (both helper streams are joined)
    for (int k = 0; k < 2; ++k) {
        if (chunk > k) MCQ_CUDA(cudaStreamWaitEvent(st, r.ev_out[k], 0));
    }
    return MCQ_OK;
}
