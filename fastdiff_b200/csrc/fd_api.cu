// C-ABI of the B200-native FastDiff sampling path (see include/fastdiff_b200.h) and the per-step
// kernel schedule.  No PyTorch types, no host synchronisation, no allocation after load_weights.
#include "../../include/fastdiff_b200.h"
#include "fd_blob.h"
#include "fd_common.cuh"
#include "fd_kernels_simt.cuh"
#include "fd_kernels_tc.cuh"   // under FD_EMU only the fp16-piece kernels (k_lvc_layer_h, k_kp_hidden_tc) compile, on the tcemu.h model

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

using namespace fd;

struct fd_handle {
    int device = 0;
    float* blob = nullptr;       // device copy of the packed weights
    size_t blob_floats = 0;
    uint64_t sec_off[FD_S_COUNT];
    uint64_t sec_cnt[FD_S_COUNT];
    float final_w[7 * C];        // host copies of tiny tensors passed by value
    float final_b = 0.f;
    int mode = FD_MODE_FP32_SIMT;
    int mode_set_by_user = 0;
    int stop_after = 99;
    int tc_upsample = 1;         // LVC-block upsample (blocks 1, 2) on tensor cores in the TC modes (option "tc_upsample")
    int tc_dblock = 1;           // DBlock 0 on tensor cores in the TC modes (option "tc_dblock")
    int tc_kp = 1;               // kernel-predictor hidden stack on tensor cores in mode tc_3xf16 (option "tc_kp")
    int emu_gemm_tc = 1;         // emulation build, mode tc_3xf16: 1 = the CTA-pair GEMM on the tcgen05 model, 0 = FFMA GEMM + k_emu_kern_to_pieces
    int emb_slots = EMB_SLOTS;   // reverse steps whose embeddings one k_embed launch computes (option "emb_slots", 1..64; tests use small values)
    int tc_b0 = 1;               // mode tc_3xf16: LVC block 0 on tensor cores (k_lvc_layer_b0h, its kernels written as fp16 pieces by the GEMM; option "tc_b0", 0 = SIMT k_lvc_layer<8>)
    int kimg_last = 0;           // the last run_denoiser wrote blocks 1, 2 in the merged-N image (fd_debug_read "kernels1/2")
    int b0_converted = 0;        // the last run_denoiser rewrote block 0's predicted kernels as fp16 pieces (fd_debug_read "kernels0")
    int b0_prefetch = 0;         // SIMT LVC kernel (block 0): bulk L2 prefetch of each warp's predicted kernels (option "b0_prefetch")
    int noise_draw_base = 0;     // device-noise mode: Philox draw number of a call's first noisy step minus one (option "noise_draw_base": callers that run
                                 // the reverse loop one fd_sample call per step -- time-shard mode -- keep the draw numbers of the single-call loop)
    long long noise_win_L = 0, noise_win_off = 0;   // Philox element window of the device-noise mode (fd_set_noise_window; time-shard mode)
    int graphs = 1;              // fd_sample in device-noise mode: capture the whole call (all N <= 64 steps) in a CUDA graph on first use and
                                 // replay it afterwards (option "graphs"; the workspace, shapes, schedule and options are the cache key)
    uint64_t epoch = 0;          // bumped by everything that changes what a captured graph would do (mode, options, weights, noise window)
    int final_stream = 0;        // option "final_w" = 1: final conv + update as the streaming warp kernel k_final_w.  Measured SLOWER than k_final (0.466 vs 0.335 ms
                                 // per N=4 call at config 2, round 2): kept as a cross-check of k_final's arithmetic, off by default
    int kc_res = 0;              // kernel_conv GEMM (tc_3xf16), option "kc_res" = 1: frame tile resident in shared memory, weights streamed, contiguous item ranges
                                 // (measured SLOWER, 0.92 vs 0.57 ms per launch: DESIGN.md 4c.14); 0 = whole-stage ring, items strided over the CTA pairs
    int kc_clusters = 0;         // test option "kc_clusters": cap on the CTA pairs of the kernel_conv GEMM (0 = one per SM pair); 1 makes one pair walk every frame tile
    int up4 = 1;                 // piece-row path: block 2 upsampling + skip by k_upsample_p4 (option "up4", 0 = k_upsample_tc<4, true>)
    int lvc_p = 1;               // mode tc_3xf16: LVC blocks 1, 2 on the piece-row protocol (k_lvc_p + k_upsample_tc<R, true>; option "lvc_p", 0 = k_lvc_layer_h)
    unsigned int* sat_flag = nullptr;   // device word, sticky: an fp16 piece saturated in a tensor-core kernel (fd_check_saturation)
    int overlap = 1;             // run the DBlock chain on an internal side stream, concurrently with embed -> kernel predictor -> GEMM
                                 // (option "overlap"; forked from / joined into the caller's stream with events inside every call)
#ifndef FD_EMU
    cudaStream_t side = nullptr;
    cudaStream_t cap = nullptr;      // capture stream of the graph path (the caller's stream may be the legacy default stream, which cannot capture)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
#endif
    int attrs_set = 0;
    uint64_t launches = 0;
#ifndef FD_EMU
    struct GraphEntry {
        void* ws; int B, Tm, n_steps, ddim, fill_xT, draw_base; uint64_t epoch; std::vector<fd_step> steps;
        cudaGraphExec_t exec; uint64_t n_launches; uint64_t last_use;
    };
    std::vector<GraphEntry> gcache;
    uint64_t gtick = 0, graph_replays = 0, graph_captures = 0;
    unsigned long long* seed_dev = nullptr;   // device word read by the captured noise kernels
#endif
    std::string err;
    void* tc_state = nullptr;    // tensor-core path resources (tensor maps etc.)
    // optional per-kernel-class device timing (bench.py's live roofline number)
    int timing = 0;
    struct TimedLaunch { int cls; cudaEvent_t a, b; };
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> ev_pool;
};

enum fd_kernel_class { KC_EMBED, KC_KP_HIDDEN, KC_KC_GEMM, KC_DBLOCK, KC_UPSAMPLE, KC_LVC0, KC_LVC1, KC_LVC2, KC_FINAL, KC_FILL, KC_COUNT };
static const char* const kKernelClassName[KC_COUNT] = {"embed", "kp_hidden", "kc_gemm", "dblock", "upsample", "lvc_layer_b0",
                                                      "lvc_layer_b1", "lvc_layer_b2", "final_update", "fill_normal"};

// Brackets the launches made in its scope with two events when timing is on (no-op otherwise).
struct ScopedTimer {
    fd_handle* h; cudaStream_t st; int idx = -1;
    ScopedTimer(fd_handle* h_, int cls, cudaStream_t st_) : h(h_), st(st_) {
#ifndef FD_EMU
        if (!h->timing) return;
        cudaEvent_t e[2];
        for (int i = 0; i < 2; ++i) {
            if (!h->ev_pool.empty()) { e[i] = h->ev_pool.back(); h->ev_pool.pop_back(); }
            else if (cudaEventCreate(&e[i]) != cudaSuccess) return;
        }
        h->timed.push_back({cls, e[0], e[1]});
        idx = (int)h->timed.size() - 1;
        cudaEventRecord(e[0], st);
#else
        (void)cls;
#endif
    }
    ~ScopedTimer() {
#ifndef FD_EMU
        if (idx >= 0) cudaEventRecord(h->timed[idx].b, st);
#endif
    }
};

static std::string g_create_err;

static int fail(fd_handle* h, int code, const char* fmt, ...) {
    char buf[640];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_err = buf;
    return code;
}

#define FD_CUDA(h, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(h, FD_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static const float* sec(const fd_handle* h, int s) { return h->blob + h->sec_off[s]; }

// ---- workspace layout (floats) ------------------------------------------------------------------
struct WsLayout {
    size_t emb, cnoise, hk, hk_hi, hk_lo, kern, d0, d1, d2, xa, xb, melc, xc, total;
};
static WsLayout ws_layout(int B, int Tm) {
    WsLayout w;
    auto al = [](size_t v) { return (v + 63) & ~(size_t)63; };  // 256-byte granules
    const size_t L = (size_t)Tm * HOP_TOTAL;
    size_t o = 0;
    w.emb = o;    o += al((size_t)EMB_SLOTS * B * EMB_OUT);          // one slot per reverse step of a chunk of the schedule
    w.cnoise = o; o += al((size_t)EMB_SLOTS * NBLK * B * COND);
    w.hk = o;     o += al((size_t)NBLK * B * (Tm + 2) * HID);   // hk, hk_hi, hk_lo are contiguous (one memset)
    w.hk_hi = o;  o += al((size_t)NBLK * B * (Tm + 2) * HID);
    w.hk_lo = o;  o += al((size_t)NBLK * B * (Tm + 2) * HID);
    w.kern = o;   o += al((size_t)NBLK * B * Tm * KCN);
    w.d0 = o;     o += al((size_t)B * (L / 4) * C);
    w.d1 = o;     o += al((size_t)B * (L / 32) * C);
    w.d2 = o;     o += al((size_t)B * Tm * C);
    w.xa = o;     o += al(lp_rows(B, (int)L) * C);   // (B,L,32) fp32 rows, or the padded piece rows of a layer input (fd_kernels_lvcp.cuh)
    w.xb = o;     o += al(lp_rows(B, (int)L) * C);
    w.melc = o;   o += al((size_t)B * COND * Tm);   // graph replay (fd_sample): the captured kernels read / write these copies, not the caller's tensors
    w.xc = o;     o += al((size_t)B * L);
    w.total = o;
    return w;
}

// ---- API -----------------------------------------------------------------------------------------
extern "C" const char* fd_version(void) { return "fastdiff_b200 0.1 (sm_100a)"; }

extern "C" const char* fd_last_error(fd_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int fd_create(const fd_config* cfg, int device, fd_handle** out) {
    if (!cfg || !out) return fail(nullptr, FD_ERR_INVALID, "fd_create: null argument");
    const bool ok = cfg->audio_channels == 1 && cfg->inner_channels == C && cfg->cond_channels == COND &&
                    cfg->n_upsample == NBLK && cfg->upsample_ratios[0] == 8 && cfg->upsample_ratios[1] == 8 &&
                    cfg->upsample_ratios[2] == 4 && cfg->lvc_layers_each_block == LAYERS && cfg->lvc_kernel_size == KS &&
                    cfg->kpnet_hidden_channels == HID && cfg->kpnet_conv_size == 3 &&
                    cfg->diffusion_step_embed_dim_in == EMB_IN && cfg->diffusion_step_embed_dim_mid == EMB_MID &&
                    cfg->diffusion_step_embed_dim_out == EMB_OUT;
    if (!ok)
        return fail(nullptr, FD_ERR_UNSUPPORTED,
                    "fd_create: only the architecture of modules/FastDiff/config/base.yaml:21-33 is built "
                    "(C=32, cond=80, ratios 8/8/4, 4 LVC layers k=3, kpnet 64/3, embed 128/512/512)");
#ifndef FD_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(nullptr, FD_ERR_CUDA, "fd_create: no usable CUDA device (index %d of %d)", device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, FD_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, FD_ERR_UNSUPPORTED, "fd_create: kernels are built for sm_100a only (device is sm_%d%d)",
                    prop.major, prop.minor);
#endif
    fd_handle* h = new fd_handle();
    h->device = device;
#ifndef FD_EMU
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete h;
        return fail(nullptr, FD_ERR_CUDA, "fd_create: creating the side stream / events failed");
    }
#endif
    if (cudaMalloc((void**)&h->sat_flag, 64) != cudaSuccess) { delete h; return fail(nullptr, FD_ERR_CUDA, "fd_create: cudaMalloc failed"); }
    cudaMemset(h->sat_flag, 0, 64);
#ifndef FD_EMU
    h->seed_dev = reinterpret_cast<unsigned long long*>(h->sat_flag + 4);
#endif
    *out = h;
    return FD_OK;
}

extern "C" void fd_destroy(fd_handle* h) {
    if (!h) return;
#ifndef FD_EMU
    cudaSetDevice(h->device);
    tc_destroy(h->tc_state);
#endif
#ifndef FD_EMU
    for (auto& t : h->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->cap) cudaStreamDestroy(h->cap);
#endif
#ifndef FD_EMU
    for (auto& g : h->gcache) cudaGraphExecDestroy(g.exec);
#endif
    if (h->blob) cudaFree(h->blob);
    if (h->sat_flag) cudaFree(h->sat_flag);
    delete h;
}

// Validates the header into `off` / `cnt` (local to the caller): a rejected blob leaves the handle exactly as it was.
static int parse_header(fd_handle* h, const uint64_t* hdr, size_t bytes, uint64_t* off, uint64_t* cnt) {
    if (bytes < 24 || hdr[0] != FD_BLOB_MAGIC) return fail(h, FD_ERR_INVALID, "weight blob: bad magic");
    if (hdr[1] != FD_BLOB_VERSION) return fail(h, FD_ERR_INVALID, "weight blob: version %ld, library expects %ld", (long)hdr[1], (long)FD_BLOB_VERSION);
    if (hdr[2] != FD_S_COUNT) return fail(h, FD_ERR_INVALID, "weight blob: %ld sections, library expects %ld", (long)hdr[2], (long)FD_S_COUNT);
    if (bytes < (3 + 2 * (size_t)FD_S_COUNT) * 8) return fail(h, FD_ERR_INVALID, "weight blob: truncated header");
    for (int s = 0; s < FD_S_COUNT; ++s) {
        off[s] = hdr[3 + 2 * s];
        cnt[s] = hdr[4 + 2 * s];
        if ((off[s] + cnt[s]) * 4 > bytes || (off[s] & 63))
            return fail(h, FD_ERR_INVALID, "weight blob: section %d out of range or misaligned", s);
    }
    // expected sizes of the sections whose shapes the kernels hard-code
    struct { int s; size_t n; } chk[] = {
        {FD_S_EMB_FREQ, 64}, {FD_S_FC1_WT, (size_t)EMB_IN * EMB_MID}, {FD_S_FC2_WT, (size_t)EMB_MID * EMB_OUT},
        {FD_S_FIRST_W, 7 * C}, {FD_S_FINAL_W, 7 * C}, {FD_S_DB0_CONV_W, 3 * KK * C}, {FD_S_LB0_UP_W, 16 * C * C},
        {FD_S_LB2_UP_W, 8 * C * C}, {FD_S_LB0_CONV_W, (size_t)LAYERS * KK * C}, {FD_S_LB0_KPIN_W, 5 * COND * HID},
        {FD_S_LB0_KPRES_W, 6 * 3 * HID * HID}, {FD_S_LB0_KC_W, (size_t)KCK * KCN}, {FD_S_LB2_KC_B, KCN},
        {FD_S_LB0_KCT_HI, (size_t)KCK * KCN}, {FD_S_LB2_KCT_LO, (size_t)KCK * KCN}};
    for (auto& c : chk)
        if (cnt[c.s] != c.n) return fail(h, FD_ERR_INVALID, "weight blob: section %d has the wrong size (%ld)", c.s, (long)cnt[c.s]);
    return FD_OK;
}

static int finish_load(fd_handle* h, const uint64_t* hdr_host) {
    (void)hdr_host;
    return FD_OK;
}

extern "C" int fd_load_weights(fd_handle* h, const void* blob_host, size_t bytes) {
    if (!h || !blob_host) return fail(h, FD_ERR_INVALID, "fd_load_weights: null argument");
    uint64_t off[FD_S_COUNT], cnt[FD_S_COUNT];
    int rc = parse_header(h, (const uint64_t*)blob_host, bytes, off, cnt);
    if (rc) return rc;
    FD_CUDA(h, cudaSetDevice(h->device));
    float* nb = nullptr;
    FD_CUDA(h, cudaMalloc((void**)&nb, bytes));
    if (cudaMemcpy(nb, blob_host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(nb); return fail(h, FD_ERR_CUDA, "fd_load_weights: copying the blob failed"); }
    if (h->blob) cudaFree(h->blob);
    h->blob = nb;
    ++h->epoch;
    memcpy(h->sec_off, off, sizeof off); memcpy(h->sec_cnt, cnt, sizeof cnt);
    h->blob_floats = bytes / 4;
    const float* hb = (const float*)blob_host;
    memcpy(h->final_w, hb + h->sec_off[FD_S_FINAL_W], sizeof h->final_w);
    h->final_b = hb[h->sec_off[FD_S_FINAL_B]];
#ifndef FD_EMU
    rc = tc_init(&h->tc_state, h->device, h->blob, h->sec_off, h->err);
    if (rc) return rc;
    if (!h->mode_set_by_user) h->mode = FD_MODE_TC_3XF16;   // default: tensor cores at fp32-level accuracy
#endif
    return finish_load(h, (const uint64_t*)blob_host);
}

extern "C" int fd_load_weights_dev(fd_handle* h, const void* blob_dev, size_t bytes, void* stream) {
    if (!h || !blob_dev) return fail(h, FD_ERR_INVALID, "fd_load_weights_dev: null argument");
    FD_CUDA(h, cudaSetDevice(h->device));
    const size_t hdr_bytes = (3 + 2 * (size_t)FD_S_COUNT) * 8;
    if (bytes < hdr_bytes) return fail(h, FD_ERR_INVALID, "weight blob: truncated header");
    std::vector<uint64_t> hdr(hdr_bytes / 8);
    cudaStream_t st = (cudaStream_t)stream;
    FD_CUDA(h, cudaMemcpyAsync(hdr.data(), blob_dev, hdr_bytes, cudaMemcpyDeviceToHost, st));
    FD_CUDA(h, cudaStreamSynchronize(st));
    uint64_t off[FD_S_COUNT], cnt[FD_S_COUNT];
    int rc = parse_header(h, hdr.data(), bytes, off, cnt);
    if (rc) return rc;
    float* nb = nullptr;
    FD_CUDA(h, cudaMalloc((void**)&nb, bytes));
    float fw[7 * C], fb = 0.f;
    if (cudaMemcpyAsync(nb, blob_dev, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(fw, (const float*)blob_dev + off[FD_S_FINAL_W], sizeof fw, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(&fb, (const float*)blob_dev + off[FD_S_FINAL_B], 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        cudaFree(nb);
        return fail(h, FD_ERR_CUDA, "fd_load_weights_dev: copying the blob failed");
    }
    if (h->blob) cudaFree(h->blob);
    h->blob = nb;
    ++h->epoch;
    memcpy(h->sec_off, off, sizeof off); memcpy(h->sec_cnt, cnt, sizeof cnt);
    memcpy(h->final_w, fw, sizeof fw); h->final_b = fb;
    h->blob_floats = bytes / 4;
#ifndef FD_EMU
    rc = tc_init(&h->tc_state, h->device, h->blob, h->sec_off, h->err);
    if (rc) return rc;
    if (!h->mode_set_by_user) h->mode = FD_MODE_TC_3XF16;
#endif
    return FD_OK;
}

extern "C" int fd_workspace_bytes(fd_handle* h, int B, int Tm, size_t* out) {
    if (!h || !out || B < 1 || Tm < 1) return fail(h, FD_ERR_INVALID, "fd_workspace_bytes: bad argument");
    *out = ws_layout(B, Tm).total * 4;
    return FD_OK;
}

extern "C" int fd_set_mode(fd_handle* h, int mode) {
    if (!h) return FD_ERR_INVALID;
    if (mode < FD_MODE_FP32_SIMT || mode > FD_MODE_TC_3XF16) return fail(h, FD_ERR_INVALID, "fd_set_mode: unknown mode %d", mode);
#ifdef FD_EMU
    if (mode != FD_MODE_FP32_SIMT && mode != FD_MODE_TC_3XF16)
        return fail(h, FD_ERR_UNSUPPORTED, "fd_set_mode: the emulation build models only the fp16-piece tensor-core kernels (tc_3xf16)");
#else
    if (mode != FD_MODE_FP32_SIMT && !tc_available(h->tc_state))
        return fail(h, FD_ERR_UNSUPPORTED, "fd_set_mode: tensor-core path unavailable (load weights first)");
#endif
    h->mode = mode;
    h->mode_set_by_user = 1;
    ++h->epoch;
    return FD_OK;
}
extern "C" int fd_get_mode(fd_handle* h) { return h ? h->mode : FD_ERR_INVALID; }

extern "C" int fd_set_noise_window(fd_handle* h, int64_t total_samples, int64_t offset) {
    if (!h || total_samples < 0 || offset < 0 || (total_samples && offset >= total_samples)) return fail(h, FD_ERR_INVALID, "fd_set_noise_window: bad window");
    h->noise_win_L = total_samples; h->noise_win_off = offset;
    ++h->epoch;
    return FD_OK;
}

extern "C" int fd_set_option(fd_handle* h, const char* key, int64_t value) {
    if (!h || !key) return FD_ERR_INVALID;
    if (!strcmp(key, "noise_draw_base")) { h->noise_draw_base = (int)value; return FD_OK; }   // part of the graph key: no epoch bump
    ++h->epoch;
    if (!strcmp(key, "graphs")) { h->graphs = (int)value; return FD_OK; }
    if (!strcmp(key, "stop_after")) { h->stop_after = (int)value; return FD_OK; }
    if (!strcmp(key, "tc_dblock")) { h->tc_dblock = (int)value; return FD_OK; }
    if (!strcmp(key, "tc_kp")) { h->tc_kp = (int)value; return FD_OK; }
    if (!strcmp(key, "overlap")) { h->overlap = (int)value; return FD_OK; }
    if (!strcmp(key, "b0_prefetch")) { h->b0_prefetch = (int)value; return FD_OK; }
    if (!strcmp(key, "tc_b0")) { h->tc_b0 = (int)value; return FD_OK; }
    if (!strcmp(key, "lvc_p")) { h->lvc_p = (int)value; return FD_OK; }
    if (!strcmp(key, "up4")) { h->up4 = (int)value; return FD_OK; }
    if (!strcmp(key, "kc_res")) { h->kc_res = (int)value; return FD_OK; }
    if (!strcmp(key, "kc_clusters")) { h->kc_clusters = (int)value; return FD_OK; }
    if (!strcmp(key, "final_w")) { h->final_stream = (int)value; return FD_OK; }
#ifndef FD_EMU
    if (!strcmp(key, "pdl")) { g_fd_pdl = value ? 1 : 0; return FD_OK; }   // programmatic dependent launch of the step's kernels (process-wide)
#endif
    if (!strcmp(key, "emu_gemm_tc")) { h->emu_gemm_tc = (int)value; return FD_OK; }
    if (!strcmp(key, "emb_slots")) {
        if (value < 1 || value > EMB_SLOTS) return fail(h, FD_ERR_INVALID, "fd_set_option: emb_slots must be in [1, %d]", EMB_SLOTS);
        h->emb_slots = (int)value; return FD_OK;
    }
    if (!strcmp(key, "tc_upsample")) { h->tc_upsample = (int)value; return FD_OK; }
#ifndef FD_EMU
    if (!strcmp(key, "lvc_swizzle")) { tc_set_lvc_swizzle(h->tc_state, (int)value); return FD_OK; }
    if (!strcmp(key, "kc_2cta")) { tc_set_kc_2cta(h->tc_state, (int)value); return FD_OK; }
    if (!strcmp(key, "lvc_exp") || !strcmp(key, "kc_exp")) {
        // masks that switch parts of the arithmetic or of the stores off to time what is left: wrong results by construction
        if (value != 0 && !getenv("FASTDIFF_B200_TIMING_EXPERIMENTS"))
            return fail(h, FD_ERR_INVALID, "fd_set_option: %s is a timing-experiment mask (wrong results by construction); set FASTDIFF_B200_TIMING_EXPERIMENTS=1 to allow it", key);
        if (key[0] == 'l') tc_set_lvc_exp(h->tc_state, (int)value); else tc_set_kc_exp(h->tc_state, (int)value);
        return FD_OK;
    }
#endif
    return fail(h, FD_ERR_INVALID, "fd_set_option: unknown key '%s'", key);
}

extern "C" uint64_t fd_launch_count(fd_handle* h) { return h ? h->launches : 0; }

extern "C" int fd_check_saturation(fd_handle* h, int* flag, int reset, void* stream) {
    if (!h || !flag) return fail(h, FD_ERR_INVALID, "fd_check_saturation: null argument");
    FD_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int v = 0;
    FD_CUDA(h, cudaMemcpyAsync(&v, h->sat_flag, 4, cudaMemcpyDeviceToHost, st));
    FD_CUDA(h, cudaStreamSynchronize(st));
    if (reset && v) FD_CUDA(h, cudaMemsetAsync(h->sat_flag, 0, 4, st));
    *flag = v ? 1 : 0;
    return FD_OK;
}

extern "C" int fd_timing_enable(fd_handle* h, int on) {
    if (!h) return FD_ERR_INVALID;
    h->timing = on ? 1 : 0;
    return FD_OK;
}

// Caller must have synchronised the stream.  Writes one JSON object {"class": {"ms": total, "n": launches}, ...}
// for the launches recorded since the last report, then recycles the events.
extern "C" int fd_timing_report(fd_handle* h, char* buf, size_t buf_bytes) {
    if (!h || !buf || buf_bytes < 64) return FD_ERR_INVALID;
    double ms[KC_COUNT] = {0};
    long cnt[KC_COUNT] = {0};
#ifndef FD_EMU
    for (auto& t : h->timed) {
        float e = 0.f;
        if (cudaEventElapsedTime(&e, t.a, t.b) == cudaSuccess) { ms[t.cls] += e; cnt[t.cls]++; }
        h->ev_pool.push_back(t.a); h->ev_pool.push_back(t.b);
    }
#endif
    h->timed.clear();
    size_t o = 0;
    o += snprintf(buf + o, buf_bytes - o, "{");
    bool first = true;
    for (int c = 0; c < KC_COUNT && o + 96 < buf_bytes; ++c) {
        if (!cnt[c]) continue;
        o += snprintf(buf + o, buf_bytes - o, "%s\"%s\": {\"ms\": %.6f, \"n\": %ld}", first ? "" : ", ", kKernelClassName[c], ms[c], cnt[c]);
        first = false;
    }
    snprintf(buf + o, buf_bytes - o, "}");
    return FD_OK;
}

// ---- kernel attribute setup (dynamic smem opt-in), once per handle --------------------------------
template <typename K>
static cudaError_t set_smem(K kern, int bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

static int setup_attrs(fd_handle* h) {
    if (h->attrs_set) return FD_OK;
    FD_CUDA(h, set_smem(k_kp_hidden, KP_SMEM_BYTES));
    FD_CUDA(h, set_smem(k_dblock<4, true>, db_smem_bytes<4>()));
    FD_CUDA(h, set_smem(k_dblock<8, false>, db_smem_bytes<8>()));
    FD_CUDA(h, set_smem(k_lvc_layer<8, 64, false>, lvc_smem_bytes<8, 64>()));
    FD_CUDA(h, set_smem(k_lvc_layer<64, 128, false>, lvc_smem_bytes<64, 128>()));
    FD_CUDA(h, set_smem(k_lvc_layer<256, 256, true>, lvc_smem_bytes<256, 256>()));
    h->attrs_set = 1;
    return FD_OK;
}

#define FD_CHECK_LAUNCH(h, name)                                                              \
    do {                                                                                      \
        cudaError_t e_ = cudaGetLastError();                                                  \
        if (e_ != cudaSuccess) return fail(h, FD_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
        h->launches++;                                                                        \
    } while (0)

// One evaluation of the denoiser, leaving h_final (B,L,32) in ws.xa.  t_dev may be null (t_scalar used).
#ifdef FD_EMU
// ---- emulation of mode tc_3xf16 (TEST INFRASTRUCTURE path of the emulation build only) -------------------------------------
// k_kp_hidden_tc and k_lvc_layer_h run on the tcemu.h model.  The CTA-pair kernel_conv GEMM is not modelled: the FFMA GEMM runs
// instead and this converter rewrites its fp32 output for blocks 1/2 into the fp16-piece image the LVC kernel expects -- an
// independent statement of the layout the GPU epilogue writes (k_kc_gemm_tc2<true>): per (frame, layer, tap) 64 rows (o) of 128 B =
// [32 i hi | 32 i lo] of w*S16_KERN, 16-byte chunk c at position c ^ (o & 7), in place of the fp32 row
// [(i/4) ^ (o & 7)][i % 4] of the same 128 bytes.
__global__ void __launch_bounds__(256) k_emu_kern_to_pieces(float* __restrict__ kern, size_t n_rows) {
    const size_t r = (size_t)blockIdx.x * 256 + threadIdx.x;   // row = ((frame * LAYERS + l) * 3 + k) * 64 + o
    if (r >= n_rows) return;
    const size_t per_frame = (size_t)LAYERS * 3 * 64;
    const size_t frame = r / per_frame, rem = r % per_frame;
    const int l = (int)(rem / 192), ko = (int)(rem % 192), o = ko & 63;
    float* row = kern + frame * KCN + (size_t)l * KPL + (size_t)ko * 32;
    float w[32];
    for (int i = 0; i < 32; ++i) w[i] = row[(((i >> 2) ^ (o & 7)) << 2) + (i & 3)];
    uint16_t* out = reinterpret_cast<uint16_t*>(row);
    for (int i = 0; i < 32; ++i) {
        uint16_t hi, lo;
        f16_split(w[i], S16_KERN, hi, lo);
        out[(((i >> 3) ^ (o & 7)) << 3) + (i & 7)] = hi;
        out[(((4 + (i >> 3)) ^ (o & 7)) << 3) + (i & 7)] = lo;
    }
}
// merged-N image (layout in k_kc_gemm_tc2, exp_mask bit 64), one thread per (frame, layer): reads the layer's 192 fp32 rows first, then writes the
// piece tiles T01 / T2 over the same 24,576 bytes (an independent statement of the layout the GPU epilogue writes)
__global__ void __launch_bounds__(64) k_emu_kern_to_pieces_m(float* __restrict__ kern, size_t n_layers) {
    const size_t r = (size_t)blockIdx.x * 64 + threadIdx.x;   // (frame * LAYERS + l)
    if (r >= n_layers) return;
    const size_t frame = r / LAYERS;
    const int l = (int)(r % LAYERS);
    float* lp = kern + frame * KCN + (size_t)l * KPL;
    static thread_local float w[3][64][32];
    for (int k = 0; k < 3; ++k)
        for (int o = 0; o < 64; ++o)
            for (int i = 0; i < 32; ++i) w[k][o][i] = lp[(k * 64 + o) * 32 + (((i >> 2) ^ (o & 7)) << 2) + (i & 3)];
    uint16_t* out = reinterpret_cast<uint16_t*>(lp);
    for (int k = 0; k < 3; ++k)
        for (int o = 0; o < 64; ++o)
            for (int i = 0; i < 32; ++i) {
                uint16_t hi, lo;
                f16_split(w[k][o][i], S16_KERN, hi, lo);
                if (k < 2) {
                    const int pos = ((((k << 2) + (i >> 3)) ^ (o & 7)) << 3) + (i & 7);
                    out[o * 64 + pos] = hi;
                    out[(64 + o) * 64 + pos] = lo;
                } else {
                    out[8192 + o * 64 + (((i >> 3) ^ (o & 7)) << 3) + (i & 7)] = hi;
                    out[8192 + o * 64 + (((4 + (i >> 3)) ^ (o & 7)) << 3) + (i & 7)] = lo;
                }
            }
}

static float emu_scale16(const fd_handle* h, int idx) { return h->blob[h->sec_off[FD_S_SCALES16] + idx]; }

static int emu_kp_hidden_tc(fd_handle* h, const float* mel, const float* cnoise, float* hk, float* hk_hi, float* hk_lo, int B, int Tm, cudaStream_t st) {
    KpTcParams p;
    for (int n = 0; n < NBLK; ++n) {
        p.w16[n] = sec(h, FD_S_LB0_KPW_F16 + n);
        p.in_b[n] = sec(h, FD_S_LB0_KPIN_B + n * FD_LB_STRIDE);
        p.res_b[n] = sec(h, FD_S_LB0_KPRES_B + n * FD_LB_STRIDE);
        for (int l = 0; l < 7; ++l) p.inv[n][l] = 1.f / (S16_HK * emu_scale16(h, 16 + 8 * n + l));
        p.inv[n][7] = 0.f;
    }
    const int total = NBLK * B * ((Tm + KT_VALID - 1) / KT_VALID);
    FD_LAUNCH(k_kp_hidden_tc, dim3(total < 8 ? total : 8), dim3(512), KT_SMEM_BYTES, st, p, mel, cnoise, hk, hk_hi, hk_lo, B, Tm);
    FD_CHECK_LAUNCH(h, "k_kp_hidden_tc");
    return FD_OK;
}

static int emu_lvc_layer_h(fd_handle* h, int blk, int layer, const float* x_in, const float* skip, const float* kern, float* x_out,
                           int B, int T, int Tm, int dil, cudaStream_t st) {
    LvcHParams hp;
    hp.cw16 = sec(h, blk == 1 ? FD_S_LB1_CONV_F16 : FD_S_LB2_CONV_F16) + (size_t)layer * (LH_CW_BYTES / 4);
    hp.conv_b = sec(h, FD_S_LB0_CONV_B + blk * FD_LB_STRIDE) + layer * C;
    hp.first_w = sec(h, FD_S_FIRST_W);
    hp.first_b = sec(h, FD_S_FIRST_B);
    const float inv_c = 1.f / (S16_ACT * emu_scale16(h, 4 + 4 * blk + layer)), inv_l = 1.f / (S16_ACT * S16_KERN);
    const int tiles = B * ((T + LT_TT - 1) / LT_TT);
    const int skip_in = (blk == 2 || layer == 0) ? 1 : 0, skip_out = (blk == 1 && layer < LAYERS - 1) ? 1 : 0;
    // a small grid on purpose: every group then walks a chunk of several tiles (carried halo rows, kernel reuse, prefetch one tile ahead)
    int grid = (tiles + 5) / 6; if (grid < 1) grid = 1; if (grid > 8) grid = 8;
    if (blk == 1) { auto k = k_lvc_layer_h<64, false, 2>;  FD_LAUNCH(k, dim3(grid), dim3(512), (lh_smem_bytes<64, false, 2>()), st, hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l, 0, skip_in, skip_out); }
    else          { auto k = k_lvc_layer_h<256, true, 2>;  FD_LAUNCH(k, dim3(grid), dim3(512), (lh_smem_bytes<256, true, 2>()), st, hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l, 0, skip_in, skip_out); }
    FD_CHECK_LAUNCH(h, "k_lvc_layer_h");
    return FD_OK;
}

// The CTA-pair kernel_conv GEMM (k_kc_gemm_tc2<true, 16>: 2-SM TMA, cta_group::2 MMA, multicast commit, remote arrives) on the model.
static int emu_kc_gemm_tc2(fd_handle* h, const float* hk_hi, const float* hk_lo, float* kern, int B, int Tm, cudaStream_t st, int b0_pieces, int kimg) {
    KcgMaps maps;
    const int res = h->kc_res;
    const uint32_t h_box_rows = res ? KC3_BROWS : 128;
    const uint64_t rows = (uint64_t)B * (Tm + 2);
    float inv[NBLK];
    for (int n = 0; n < NBLK; ++n) {
        const float* w16 = (n == 0 && b0_pieces) ? sec(h, FD_S_LB0_KCT_F16P) : sec(h, FD_S_LB0_KCT_F16 + n);
        emu_make_map_2d(&maps.w_hi[n], w16, KCK / 2, KCN, KCK * 2, KCG_KATOM, KCG_BM);
        emu_make_map_2d(&maps.w_lo[n], w16 + (size_t)KCN * (KCK / 2), KCK / 2, KCN, KCK * 2, KCG_KATOM, KCG_BM);
        emu_make_map_2d(&maps.h_hi[n], hk_hi + (size_t)n * rows * (HID / 2), HID / 2, rows, HID * 2, KCG_KATOM, h_box_rows);
        emu_make_map_2d(&maps.h_lo[n], hk_lo + (size_t)n * rows * (HID / 2), HID / 2, rows, HID * 2, KCG_KATOM, h_box_rows);
        inv[n] = 1.f / (emu_scale16(h, n) * S16_HK);
    }
    const int M = B * (Tm + 2) - 2;
    const int items = NBLK * (KCN / 256) * ((M + 255) / 256);
    const int cap = h->kc_clusters > 0 ? h->kc_clusters : 8;
    const int clusters = items < cap ? items : cap;
    if (b0_pieces && res) {
        auto k = k_kc_gemm_tc2<true, 16, true, true>;
        FD_LAUNCH_CLUSTER2(k, dim3(2 * clusters), dim3(64 + 32 * 16), KC3_SMEM_BYTES, st, maps, sec(h, FD_S_LB0_KC_BP), sec(h, FD_S_LB1_KC_B),
                           sec(h, FD_S_LB2_KC_B), kern, B, Tm, 1, inv[0], inv[1], inv[2], kimg ? 64 : 0);
    } else if (b0_pieces) {
        auto k = k_kc_gemm_tc2<true, 16, true>;
        FD_LAUNCH_CLUSTER2(k, dim3(2 * clusters), dim3(64 + 32 * 16), KC2_SMEM_BYTES, st, maps, sec(h, FD_S_LB0_KC_BP), sec(h, FD_S_LB1_KC_B),
                           sec(h, FD_S_LB2_KC_B), kern, B, Tm, 1, inv[0], inv[1], inv[2], kimg ? 64 : 0);
    } else if (res) {
        auto k = k_kc_gemm_tc2<true, 16, false, true>;
        FD_LAUNCH_CLUSTER2(k, dim3(2 * clusters), dim3(64 + 32 * 16), KC3_SMEM_BYTES, st, maps, sec(h, FD_S_LB0_KC_B), sec(h, FD_S_LB1_KC_B),
                           sec(h, FD_S_LB2_KC_B), kern, B, Tm, 1, inv[0], inv[1], inv[2], kimg ? 64 : 0);
    } else {
        auto k = k_kc_gemm_tc2<true, 16>;
        FD_LAUNCH_CLUSTER2(k, dim3(2 * clusters), dim3(64 + 32 * 16), KC2_SMEM_BYTES, st, maps, sec(h, FD_S_LB0_KC_B), sec(h, FD_S_LB1_KC_B),
                           sec(h, FD_S_LB2_KC_B), kern, B, Tm, 1, inv[0], inv[1], inv[2], kimg ? 64 : 0);
    }
    FD_CHECK_LAUNCH(h, "k_kc_gemm_tc2");
    return FD_OK;
}

// DBlock 0 and the upsampling of blocks 1/2 still run the 3xTF32 kernels in mode tc_3xf16: modelled as well
static int emu_dblock0_tc(fd_handle* h, const float* audio, float* d0, int B, int L, cudaStream_t st) {
    DbTcParams p;
    p.cw_hi = sec(h, FD_S_DB0_CONVT_HI); p.cw_lo = sec(h, FD_S_DB0_CONVT_LO);
    p.rw_hi = sec(h, FD_S_DB0_REST_HI);  p.rw_lo = sec(h, FD_S_DB0_REST_LO);
    p.conv_b = sec(h, FD_S_DB0_CONV_B);  p.res_b = sec(h, FD_S_DB0_RES_B);
    p.first_w = sec(h, FD_S_FIRST_W);    p.first_b = sec(h, FD_S_FIRST_B);
    const int To = L / 4, total = B * ((To + DT_VALID - 1) / DT_VALID);
    FD_LAUNCH(k_dblock0_tc, dim3(total < 6 ? total : 6), dim3(512), DT_SMEM_BYTES, st, p, audio, d0, B, L, To, 1);
    FD_CHECK_LAUNCH(h, "k_dblock0_tc");
    return FD_OK;
}

static int emu_upsample_tc(fd_handle* h, int blk, const float* in, float* out, int B, int Tin, cudaStream_t st) {
    const float* wh = sec(h, blk == 1 ? FD_S_LB1_UPT_HI : FD_S_LB2_UPT_HI);
    const float* wl = sec(h, blk == 1 ? FD_S_LB1_UPT_LO : FD_S_LB2_UPT_LO);
    const float* bias = sec(h, FD_S_LB0_UP_B + blk * FD_LB_STRIDE);
    const int total = B * ((Tin + 127) / 128);
    const int grid = total < 6 ? total : 6;
    if (blk == 1) { auto k = k_upsample_tc<8>; FD_LAUNCH(k, dim3(grid), dim3(512), (ut_smem_bytes<8>()), st, wh, wl, bias, in, out, B, Tin, 1, UpPOut()); }
    else          { auto k = k_upsample_tc<4>; FD_LAUNCH(k, dim3(grid), dim3(512), (ut_smem_bytes<4>()), st, wh, wl, bias, in, out, B, Tin, 1, UpPOut()); }
    FD_CHECK_LAUNCH(h, "k_upsample_tc");
    return FD_OK;
}
static int emu_lvc_layer_b0(fd_handle* h, int layer, const float* x_in, const float* skip, const float* kern, float* x_out,
                            int B, int T, int Tm, int dil, cudaStream_t st) {
    LvcHParams hp;
    hp.cw16 = sec(h, FD_S_LB0_CONV_F16) + (size_t)layer * (LH_CW_BYTES / 4);
    hp.conv_b = sec(h, FD_S_LB0_CONV_B) + layer * C;
    hp.first_w = nullptr; hp.first_b = nullptr;
    const float inv_c = 1.f / (S16_ACT * emu_scale16(h, 4 + layer)), inv_l = 1.f / (S16_ACT * S16_KERN);
    const int tiles = B * ((T + LT_TT - 1) / LT_TT);
    int grid = (tiles + 2) / 3; if (grid < 1) grid = 1; if (grid > 8) grid = 8;   // several tiles per CTA: the ring runs across tiles
    FD_LAUNCH(k_lvc_layer_b0h, dim3(grid), dim3(512), LB0_SMEM_BYTES, st, hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l,
              layer == 0 ? 1 : 0, layer < LAYERS - 1 ? 1 : 0);
    FD_CHECK_LAUNCH(h, "k_lvc_layer_b0h");
    return FD_OK;
}
// piece-row protocol (fd_kernels_lvcp.cuh) on the model
static int emu_zero_pads(fd_handle* h, float* buf, int B, int T, cudaStream_t st) {
    FD_LAUNCH(k_zero_pads, dim3(B + 1), dim3(256), 0, st, buf, B, T);
    FD_CHECK_LAUNCH(h, "k_zero_pads");
    return FD_OK;
}
static int emu_upsample_p(fd_handle* h, int blk, const float* in, const float* skip, float* p_out, int B, int Tin, cudaStream_t st) {
    const float* wh = sec(h, blk == 1 ? FD_S_LB1_UPT_HI : FD_S_LB2_UPT_HI);
    const float* wl = sec(h, blk == 1 ? FD_S_LB1_UPT_LO : FD_S_LB2_UPT_LO);
    const float* bias = sec(h, FD_S_LB0_UP_B + blk * FD_LB_STRIDE);
    UpPOut po;
    po.skip = skip; po.first_w = sec(h, FD_S_FIRST_W); po.first_b = sec(h, FD_S_FIRST_B); po.p_out = p_out; po.sat = h->sat_flag;
    const int total = B * ((Tin + 127) / 128);
    const int grid = total < 6 ? total : 6;
    if (blk == 1) { auto k = k_upsample_tc<8, true>; FD_LAUNCH(k, dim3(grid), dim3(512), (ut_smem_bytes<8>() + UT_PEXTRA), st, wh, wl, bias, in, p_out, B, Tin, 1, po); }
    else          { auto k = k_upsample_tc<4, true>; FD_LAUNCH(k, dim3(grid), dim3(512), (ut_smem_bytes<4>() + UT_PEXTRA), st, wh, wl, bias, in, p_out, B, Tin, 1, po); }
    FD_CHECK_LAUNCH(h, "k_upsample_tc<POUT>");
    return FD_OK;
}
static int emu_upsample_p4(fd_handle* h, const float* in, const float* audio, float* p_out, int B, int Tin, cudaStream_t st) {
    Up4Params p;
    p.w16 = sec(h, FD_S_LB2_UP_F16M); p.first16u = sec(h, FD_S_FIRST_F16U); p.bias = sec(h, FD_S_LB0_UP_B + 2 * FD_LB_STRIDE);
    p.in = in; p.audio = audio; p.p_out = p_out; p.sat = h->sat_flag; p.B = B; p.Tin = Tin;
    p.inv = 1.f / (S16_ACT * emu_scale16(h, 41));
    const int total = B * ((Tin + 127) / 128);
    FD_LAUNCH(k_upsample_p4, dim3(total < 6 ? total : 6), dim3(512), U4_SMEM_BYTES, st, p);
    FD_CHECK_LAUNCH(h, "k_upsample_p4");
    return FD_OK;
}
static int emu_lvc_p_layer(fd_handle* h, int blk, int layer, const float* p_in, const float* skip, const float* kern, float* p_out, float* f_out,
                           int B, int T, int Tm, int dil, cudaStream_t st) {
    LvcPParams p;
    p.cw16 = sec(h, blk == 1 ? FD_S_LB1_CONV_F16M : FD_S_LB2_CONV_F16M) + (size_t)layer * (LP_CW_BYTES / 4);
    p.conv_b = sec(h, FD_S_LB0_CONV_B + blk * FD_LB_STRIDE) + layer * C;
    p.first16 = sec(h, FD_S_FIRST_F16);
    p.p_in = p_in; p.skip = skip; p.kern = kern; p.p_out = p_out; p.f_out = f_out; p.sat = h->sat_flag;
    p.B = B; p.T = T; p.Tm = Tm; p.dil = dil;
    p.inv_c = 1.f / (S16_ACT * emu_scale16(h, 4 + 4 * blk + layer));
    p.inv_l = 1.f / (S16_ACT * S16_KERN);
    p.inv_sk = 1.f / (LP_S_AU * emu_scale16(h, 40));
    const int tiles = B * ((T + LP_TT - 1) / LP_TT);
    int grid = (tiles + 3) / 4; if (grid < 1) grid = 1; if (grid > 8) grid = 8;   // a few tiles per CTA: carried rows, kernel reuse, every ring wraps
    if (blk == 1) { auto k = k_lvc_p<64>;  FD_LAUNCH(k, dim3(grid), dim3(LP_THREADS), (lp_smem_bytes<64>()), st, p); }
    else          { auto k = k_lvc_p<256>; FD_LAUNCH(k, dim3(grid), dim3(LP_THREADS), (lp_smem_bytes<256>()), st, p); }
    FD_CHECK_LAUNCH(h, "k_lvc_p");
    return FD_OK;
}
#endif  // FD_EMU

static int launch_embed(fd_handle* h, const float* t_dev, const EmbedSteps& ts, int nslots, int B, float* ws, cudaStream_t st) {
    const WsLayout w = ws_layout(B, 1);   // emb / cnoise sit at the front of the workspace: their offsets do not depend on T'
    EmbedParams p;
    p.freq = sec(h, FD_S_EMB_FREQ);
    p.w1t = sec(h, FD_S_FC1_WT); p.b1 = sec(h, FD_S_FC1_B);
    p.w2t = sec(h, FD_S_FC2_WT); p.b2 = sec(h, FD_S_FC2_B);
    for (int n = 0; n < NBLK; ++n) {
        p.fct_wt[n] = sec(h, FD_S_LB0_FCT_WT + n * FD_LB_STRIDE);
        p.fct_b[n] = sec(h, FD_S_LB0_FCT_B + n * FD_LB_STRIDE);
    }
    ScopedTimer tm(h, KC_EMBED, st);
    FD_LAUNCH_PDL(k_embed, dim3(B, nslots), dim3(512), 0, st, p, t_dev, ts, ws + w.emb, ws + w.cnoise, B);
    FD_CHECK_LAUNCH(h, "k_embed");
    return FD_OK;
}

// emb_slot >= 0: the step embedding of this evaluation is already in that slot (fd_sample computes a chunk of steps per launch);
// emb_slot < 0: compute it here into slot 0 from t_dev / t_scalar.
static int run_denoiser(fd_handle* h, const float* x_dev, const float* mel_dev, const float* t_dev, float t_scalar,
                        int B, int Tm, float* ws, cudaStream_t st, int emb_slot = -1) {
    const WsLayout w = ws_layout(B, Tm);
    const int L = Tm * HOP_TOTAL;
    const int slot = emb_slot < 0 ? 0 : emb_slot;
    float* cnoise = ws + w.cnoise + (size_t)slot * NBLK * B * COND; float* hk = ws + w.hk; float* kern = ws + w.kern;
    float* d0 = ws + w.d0; float* d1 = ws + w.d1; float* d2 = ws + w.d2; float* xa = ws + w.xa; float* xb = ws + w.xb;
    h->b0_converted = 0;

    // -- first_audio_conv + the three DiffusionDBlocks: independent of the kernel-predictor path, so they run on the side stream
    //    (forked here, joined before the first LVC block) unless a debugging stop or the option "overlap" = 0 asks for serial order
    bool up0_done = false;
    auto run_dblocks = [&](cudaStream_t sd) -> int {
        DbParams p;
        p.first_w = sec(h, FD_S_FIRST_W); p.first_b = sec(h, FD_S_FIRST_B);
        const float* ins[3] = {x_dev, d0, d1};
        float* outs[3] = {d0, d1, d2};
        const int tin[3] = {L, L / 4, L / 32}, tout[3] = {L / 4, L / 32, L / 256};
        for (int n = 0; n < NBLK; ++n) {
            p.res_w = sec(h, FD_S_DB0_RES_W + n * FD_DB_STRIDE);   p.res_b = sec(h, FD_S_DB0_RES_B + n * FD_DB_STRIDE);
            p.conv_w = sec(h, FD_S_DB0_CONV_W + n * FD_DB_STRIDE); p.conv_b = sec(h, FD_S_DB0_CONV_B + n * FD_DB_STRIDE);
            const dim3 grid((tout[n] + DB_TO - 1) / DB_TO, B);
            ScopedTimer tm(h, KC_DBLOCK, sd);
            bool done_tc = false;
#ifdef FD_EMU
            if (n == 0 && h->mode == FD_MODE_TC_3XF16 && h->tc_dblock) {
                int rc = emu_dblock0_tc(h, x_dev, d0, B, L, sd);
                if (rc) return rc;
                done_tc = true;
            }
#else
            if (n == 0 && h->mode != FD_MODE_FP32_SIMT && h->tc_dblock) {
                int rc = tc_dblock0(h->tc_state, h->mode, x_dev, d0, B, L, sd, h->err, &h->launches);
                if (rc) return rc;
                done_tc = true;
            }
#endif
            if (done_tc) continue;
            if (n == 0) { auto k = k_dblock<4, true>;  FD_LAUNCH_PDL(k, grid, dim3(256), db_smem_bytes<4>(), sd, p, ins[n], outs[n], tin[n], tout[n]); }
            else        { auto k = k_dblock<8, false>; FD_LAUNCH_PDL(k, grid, dim3(256), db_smem_bytes<8>(), sd, p, ins[n], outs[n], tin[n], tout[n]); }
            FD_CHECK_LAUNCH(h, "k_dblock");
        }
        // LVC block 0's upsampling reads the last DBlock output only: it rides on the same stream (off the critical path when the chain is
        // forked to the side stream: embed -> kernel predictor -> GEMM take longer than DBlocks + this launch)
        if (h->stop_after > 2) {
            ScopedTimer tm(h, KC_UPSAMPLE, sd);
            auto k = k_upsample<8>;
            FD_LAUNCH_PDL(k, dim3((Tm + 31) / 32, B), dim3(256), 0, sd, sec(h, FD_S_LB0_UP_W), sec(h, FD_S_LB0_UP_B), d2, xa, Tm);
            FD_CHECK_LAUNCH(h, "k_upsample");
            up0_done = true;
        }
        return FD_OK;
    };
    bool forked = false;
#ifndef FD_EMU
    if (h->overlap && h->side && h->stop_after > 2) {
        FD_CUDA(h, cudaEventRecord(h->ev_fork, st));
        FD_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        int rc = run_dblocks(h->side);
        if (rc) return rc;
        forked = true;
    }
#endif
    // -- step embedding + per-block condition offsets
    if (emb_slot < 0) {
        EmbedSteps ts;
        ts.t[0] = t_scalar;
        int rc = launch_embed(h, t_dev, ts, 1, B, ws, st);
        if (rc) return rc;
    }
    // -- kernel predictor: hidden stack, then kernel_conv+bias_conv GEMM (all three blocks per launch)
    {
        KpParams p;
        for (int n = 0; n < NBLK; ++n) {
            p.in_w[n] = sec(h, FD_S_LB0_KPIN_W + n * FD_LB_STRIDE);   p.in_b[n] = sec(h, FD_S_LB0_KPIN_B + n * FD_LB_STRIDE);
            p.res_w[n] = sec(h, FD_S_LB0_KPRES_W + n * FD_LB_STRIDE); p.res_b[n] = sec(h, FD_S_LB0_KPRES_B + n * FD_LB_STRIDE);
        }
        ScopedTimer tm(h, KC_KP_HIDDEN, st);
        bool kp_done = false;
#ifdef FD_EMU
        if (h->mode == FD_MODE_TC_3XF16 && h->tc_kp) {
            int rc = emu_kp_hidden_tc(h, mel_dev, cnoise, hk, ws + w.hk_hi, ws + w.hk_lo, B, Tm, st);
            if (rc) return rc;
            kp_done = true;
        }
#else
        if (h->mode == FD_MODE_TC_3XF16 && h->tc_kp) {
            int rc = tc_kp_hidden(h->tc_state, mel_dev, cnoise, hk, ws + w.hk_hi, ws + w.hk_lo, B, Tm, st, h->err, &h->launches);
            if (rc) return rc;
            kp_done = true;
        }
#endif
        if (!kp_done) {
        FD_LAUNCH(k_kp_hidden, dim3((Tm + KP_FT - 1) / KP_FT, B, NBLK), dim3(256), KP_SMEM_BYTES, st, p, mel_dev, cnoise, hk, ws + w.hk_hi, ws + w.hk_lo, B, Tm,
                  h->mode == FD_MODE_TC_3XF16 ? 1 : 0);
        FD_CHECK_LAUNCH(h, "k_kp_hidden");
        }
    }
#ifdef FD_EMU
    const bool simt_gemm = !(h->mode == FD_MODE_TC_3XF16 && h->emu_gemm_tc);
#else
    const bool simt_gemm = h->mode == FD_MODE_FP32_SIMT;
#endif
    // LVC block 0 on tensor cores (mode tc_3xf16, option tc_b0): the GEMM writes the block's kernels as fp16 pieces itself
    const bool b0_tc = h->tc_b0 && h->mode == FD_MODE_TC_3XF16 && !simt_gemm;
    const bool b0_gemm_pieces = b0_tc;
    // blocks 1, 2 consumed by k_lvc_p: the GEMM writes their predicted kernels in the merged-N (SWIZZLE_64B, piece-major) operand image
    const int kimg = (h->mode == FD_MODE_TC_3XF16 && h->lvc_p && h->tc_upsample) ? 1 : 0;
    h->kimg_last = kimg;
#ifdef FD_EMU
    if (!simt_gemm) {
        ScopedTimer tm(h, KC_KC_GEMM, st);
        int rc = emu_kc_gemm_tc2(h, ws + w.hk_hi, ws + w.hk_lo, kern, B, Tm, st, b0_gemm_pieces ? 1 : 0, kimg);
        if (rc) return rc;
    }
#endif
    if (simt_gemm) {
        KcParams p;
        for (int n = 0; n < NBLK; ++n) { p.w[n] = sec(h, FD_S_LB0_KC_W + n * FD_LB_STRIDE); p.b[n] = sec(h, FD_S_LB0_KC_B + n * FD_LB_STRIDE); }
        const int M = B * (Tm + 2) - 2;
        ScopedTimer tm(h, KC_KC_GEMM, st);
        FD_LAUNCH(k_kc_gemm_simt, dim3(KCN / 128, (M + 127) / 128, NBLK), dim3(256), 0, st, p, hk, kern, B, Tm);
        FD_CHECK_LAUNCH(h, "k_kc_gemm_simt");
#ifdef FD_EMU
        if (h->mode == FD_MODE_TC_3XF16) {   // blocks 1, 2: fp32 image -> fp16-piece image (see k_emu_kern_to_pieces)
            const size_t n_rows = (size_t)2 * B * Tm * LAYERS * 3 * 64;
            if (kimg) {   // merged-N image of k_lvc_p
                const size_t n_layers = (size_t)2 * B * Tm * LAYERS;
                FD_LAUNCH(k_emu_kern_to_pieces_m, dim3((unsigned)((n_layers + 63) / 64)), dim3(64), 0, st, kern + (size_t)B * Tm * KCN, n_layers);
            } else {
                FD_LAUNCH(k_emu_kern_to_pieces, dim3((unsigned)((n_rows + 255) / 256)), dim3(256), 0, st, kern + (size_t)B * Tm * KCN, n_rows);
            }
            FD_CHECK_LAUNCH(h, "k_emu_kern_to_pieces");
        }
#endif
    } else {
#ifndef FD_EMU
        ScopedTimer tm(h, KC_KC_GEMM, st);
        int rc = tc_kc_gemm(h->tc_state, h->mode, ws + w.hk_hi, ws + w.hk_lo, kern, B, Tm, st, h->err, &h->launches, b0_gemm_pieces ? 1 : 0, kimg, h->kc_res, h->kc_clusters);
        if (rc) return rc;
#endif
    }
    if (b0_gemm_pieces) h->b0_converted = 1;
    if (h->stop_after <= 1) return FD_OK;

    if (!forked) { int rc = run_dblocks(st); if (rc) return rc; }
#ifndef FD_EMU
    if (forked) {
        FD_CUDA(h, cudaEventRecord(h->ev_join, h->side));
        FD_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    }
#endif
    if (h->stop_after <= 2) return FD_OK;

    // -- the three TimeAware_LVCBlocks
    const float* blk_in = d2;
    int Tin = Tm;
    float* cur = xa; float* oth = xb;
    for (int n = 0; n < NBLK; ++n) {
        const int r = ratio_of(n), T = Tin * r;
        const float* upw = sec(h, FD_S_LB0_UP_W + n * FD_LB_STRIDE);
        const float* upb = sec(h, FD_S_LB0_UP_B + n * FD_LB_STRIDE);
        const float* skip = (n == 0) ? d1 : (n == 1 ? d0 : x_dev);
        // default path of mode tc_3xf16 for blocks 1, 2: the state travels between layers as fp16 operand pieces (fd_kernels_lvcp.cuh)
        const bool use_p = (n >= 1 && h->mode == FD_MODE_TC_3XF16 && h->lvc_p && h->tc_upsample);
        if (use_p) {   // blk_in lives in `oth`: zero the pads of `cur`, up-sample into it, then zero the pads of `oth` for layer 0's output
            ScopedTimer tm(h, KC_UPSAMPLE, st);
#ifdef FD_EMU
            int rc = emu_zero_pads(h, cur, B, T, st);
            if (!rc) rc = (n == 2 && h->up4) ? emu_upsample_p4(h, blk_in, skip, cur, B, Tin, st) : emu_upsample_p(h, n, blk_in, skip, cur, B, Tin, st);
            if (!rc) rc = emu_zero_pads(h, oth, B, T, st);
#else
            int rc = tc_zero_pads(cur, B, T, st, h->err, &h->launches);
            if (!rc) rc = (n == 2 && h->up4) ? tc_upsample_p4(h->tc_state, blk_in, skip, cur, h->sat_flag, B, Tin, st, h->err, &h->launches)
                                             : tc_upsample_p(h->tc_state, n, blk_in, skip, cur, h->sat_flag, B, Tin, st, h->err, &h->launches);
            if (!rc) rc = tc_zero_pads(oth, B, T, st, h->err, &h->launches);
#endif
            if (rc) return rc;
        } else
        {
            const dim3 grid((Tin + 31) / 32, B);
            ScopedTimer tm(h, KC_UPSAMPLE, st);
            bool up_done = false;
#ifdef FD_EMU
            if (n >= 1 && h->mode == FD_MODE_TC_3XF16 && h->tc_upsample) {
                int rc = emu_upsample_tc(h, n, blk_in, cur, B, Tin, st);
                if (rc) return rc;
                up_done = true;
            }
#else
            if (n >= 1 && h->mode != FD_MODE_FP32_SIMT && h->tc_upsample) {
                int rc = tc_upsample(h->tc_state, h->mode, n, blk_in, cur, B, Tin, st, h->err, &h->launches);
                if (rc) return rc;
                up_done = true;
            }
#endif
            if (n == 0 && up0_done) up_done = true;   // launched with the DBlock chain
            if (!up_done) {
                if (r == 8) { auto k = k_upsample<8>; FD_LAUNCH_PDL(k, grid, dim3(256), 0, st, upw, upb, blk_in, cur, Tin); }
                else        { auto k = k_upsample<4>; FD_LAUNCH_PDL(k, grid, dim3(256), 0, st, upw, upb, blk_in, cur, Tin); }
                FD_CHECK_LAUNCH(h, "k_upsample");
            }
        }
        const float* kern_n = kern + (size_t)n * B * Tm * KCN;
        const bool b0_here = (n == 0 && b0_tc);
        for (int i = 0; i < LAYERS; ++i) {
            LvcParams p;
            p.conv_w = sec(h, FD_S_LB0_CONV_W + n * FD_LB_STRIDE) + i * KK * C;
            p.conv_b = sec(h, FD_S_LB0_CONV_B + n * FD_LB_STRIDE) + i * C;
            p.first_w = sec(h, FD_S_FIRST_W); p.first_b = sec(h, FD_S_FIRST_B);
            int dil = 1; for (int q = 0; q < i; ++q) dil *= 3;
            const float* kl = kern_n + i * KPL;
            bool done = false;
            ScopedTimer tm(h, KC_LVC0 + n, st);
            if (use_p) {
                float* p_out = i < LAYERS - 1 ? oth : nullptr;
                float* f_out = i < LAYERS - 1 ? nullptr : oth;
#ifdef FD_EMU
                int rc = emu_lvc_p_layer(h, n, i, cur, skip, kl, p_out, f_out, B, T, Tm, dil, st);
#else
                int rc = tc_lvc_p_layer(h->tc_state, n, i, cur, skip, kl, p_out, f_out, h->sat_flag, B, T, Tm, dil, st, h->err, &h->launches);
#endif
                if (rc) return rc;
                done = true;
            }
#ifdef FD_EMU
            else if (b0_here) {
                int rc = emu_lvc_layer_b0(h, i, cur, skip, kl, oth, B, T, Tm, dil, st);
                if (rc) return rc;
                done = true;
            } else if (h->mode == FD_MODE_TC_3XF16 && n >= 1) {
                int rc = emu_lvc_layer_h(h, n, i, cur, skip, kl, oth, B, T, Tm, dil, st);
                if (rc) return rc;
                done = true;
            }
#else
            else if (b0_here) {
                int rc = tc_lvc_layer_b0(h->tc_state, i, cur, skip, kl, oth, B, T, Tm, dil, st, h->err, &h->launches);
                if (rc) return rc;
                done = true;
            } else if (h->mode != FD_MODE_FP32_SIMT) {
                int rc = tc_lvc_layer(h->tc_state, h->mode, n, i, cur, skip, kl, oth, B, T, Tm, dil, st, h->err, &h->launches, &done);
                if (rc) return rc;
            }
#endif
            if (!done) {
                if (n == 0)      { auto k = k_lvc_layer<8, 64, false>;    FD_LAUNCH(k, dim3((T + 63) / 64, B), dim3(256), (lvc_smem_bytes<8, 64>()), st, p, cur, skip, kl, oth, T, Tm, dil, h->b0_prefetch); }
                else if (n == 1) { auto k = k_lvc_layer<64, 128, false>;  FD_LAUNCH(k, dim3((T + 127) / 128, B), dim3(256), (lvc_smem_bytes<64, 128>()), st, p, cur, skip, kl, oth, T, Tm, dil, h->b0_prefetch); }
                else             { auto k = k_lvc_layer<256, 256, true>;  FD_LAUNCH(k, dim3((T + 255) / 256, B), dim3(256), (lvc_smem_bytes<256, 256>()), st, p, cur, skip, kl, oth, T, Tm, dil, h->b0_prefetch); }
                FD_CHECK_LAUNCH(h, "k_lvc_layer");
            }
            float* tmp = cur; cur = oth; oth = tmp;
        }
        if (h->stop_after <= 3 + n) return FD_OK;
        blk_in = cur;   // block output lives in `cur`; the next block upsamples it into the other buffer
        Tin = T;
        float* tmp = cur; cur = oth; oth = tmp;
    }
    return FD_OK;
}

// After run_denoiser the last block's output is in xa (up -> xa; 4 layers: xa->xb->xa->xb->xa ... per block the
// parity of buffer swaps is fixed, see final_buffer()).
static float* final_buffer(float* ws, int B, int Tm) {
    // block 0: up->xa, layers end in xa; swap -> cur=xb. block 1: up->xb, ends in xb; swap -> cur=xa.
    // block 2: up->xa, ends in xa.
    return ws + ws_layout(B, Tm).xa;
}
static float* block_out_buffer(float* ws, int B, int Tm, int n) {
    const WsLayout w = ws_layout(B, Tm);
    return ws + (n == 1 ? w.xb : w.xa);
}

static int check_args(fd_handle* h, const void* a, const void* b, const void* c, int B, int Tm, void* ws, size_t ws_bytes) {
    if (!h) return FD_ERR_INVALID;
    if (!h->blob) return fail(h, FD_ERR_STATE, "weights not loaded (call fd_load_weights first)");
    if (!a || !b || !c || !ws) return fail(h, FD_ERR_INVALID, "null device pointer");
    if (B < 1 || Tm < 1) return fail(h, FD_ERR_INVALID, "bad shape: B=%d T'=%d", B, Tm);
    if ((size_t)B * Tm * HOP_TOTAL > 0x7fffffffULL / 4) return fail(h, FD_ERR_INVALID, "B*L too large for 32-bit indexing of one call; split the batch");
    if (ws_bytes < ws_layout(B, Tm).total * 4) return fail(h, FD_ERR_INVALID, "workspace too small: %ld bytes given, %ld needed", (long)ws_bytes, (long)(ws_layout(B, Tm).total * 4));
    return FD_OK;
}

static void fill_final(const fd_handle* h, FinalParams& fp) {
    memcpy(fp.w, h->final_w, sizeof fp.w);
    fp.b = h->final_b;
    fp.mode = 0; fp.coef = 0; fp.div = 1; fp.sigma = 0; fp.c1 = fp.c2 = fp.c3 = 0; fp.add_noise = 0; fp.draw = 0; fp.seed = 0;
    fp.win.win_L = h->noise_win_L; fp.win.win_off = h->noise_win_off; fp.seed_ptr = nullptr;
}

extern "C" int fd_denoise(fd_handle* h, const float* x_dev, const float* mel_dev, const float* t_dev, float* eps_dev,
                          int B, int Tm, void* workspace_dev, size_t workspace_bytes, void* stream) {
    int rc = check_args(h, x_dev, mel_dev, t_dev, B, Tm, workspace_dev, workspace_bytes);
    if (rc) return rc;
    if (!eps_dev) return fail(h, FD_ERR_INVALID, "null device pointer");
    FD_CUDA(h, cudaSetDevice(h->device));
    rc = setup_attrs(h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const WsLayout w = ws_layout(B, Tm);
    FD_CUDA(h, cudaMemsetAsync(ws + w.hk, 0, (w.kern - w.hk) * 4, st));
    rc = run_denoiser(h, x_dev, mel_dev, t_dev, 0.f, B, Tm, ws, st);
    if (rc) return rc;
    if (h->stop_after < 6) return FD_OK;
    FinalParams fp;
    fill_final(h, fp);
    const int L = Tm * HOP_TOTAL;
    {
        ScopedTimer tm(h, KC_FINAL, st);
        if (h->final_stream) FD_LAUNCH_PDL(k_final_w, dim3(L / 256, B), dim3(256), 0, st, fp, final_buffer(ws, B, Tm), x_dev, (const float*)nullptr, eps_dev, (float*)nullptr, L);
        else            FD_LAUNCH_PDL(k_final, dim3(L / 256, B), dim3(256), 0, st, fp, final_buffer(ws, B, Tm), x_dev, (const float*)nullptr, eps_dev, (float*)nullptr, L);
        FD_CHECK_LAUNCH(h, "k_final");
    }
    return FD_OK;
}

// The N-step loop of util.py:216-234 as a sequence of launches on `st` (no host synchronisation).  seed_ptr != nullptr: the noise kernels
// read the seed from that device word (graph replay).
static int sample_body(fd_handle* h, float* x_dev, const float* mel_dev, const fd_step* steps, int n_steps, const float* noise_dev, uint64_t seed,
                       const unsigned long long* seed_ptr, int fill_xT, int ddim, float* seq_dev, int B, int Tm, float* ws, cudaStream_t st) {
    const WsLayout w = ws_layout(B, Tm);
    const int L = Tm * HOP_TOTAL;
    const size_t n = (size_t)B * L;
    int rc;
    FD_CUDA(h, cudaMemsetAsync(ws + w.hk, 0, (w.kern - w.hk) * 4, st));
    NoiseWin win; win.win_L = h->noise_win_L; win.win_off = h->noise_win_off;
    if (fill_xT) {
        ScopedTimer tm(h, KC_FILL, st);
        FD_LAUNCH_PDL(k_fill_normal, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x_dev, L, n, seed, 0u, win, seed_ptr);
        FD_CHECK_LAUNCH(h, "k_fill_normal");
    }
    if (seq_dev) FD_CUDA(h, cudaMemcpyAsync(seq_dev, x_dev, n * 4, cudaMemcpyDeviceToDevice, st));
    int draw = 0;
    for (int i = 0; i < n_steps; ++i) {
        if (i % h->emb_slots == 0) {   // step embeddings of the next (up to) 64 reverse steps in one launch
            EmbedSteps ts;
            const int ns = n_steps - i < h->emb_slots ? n_steps - i : h->emb_slots;
            for (int s2 = 0; s2 < ns; ++s2) ts.t[s2] = steps[i + s2].t;
            rc = launch_embed(h, nullptr, ts, ns, B, ws, st);
            if (rc) return rc;
        }
        rc = run_denoiser(h, x_dev, mel_dev, nullptr, steps[i].t, B, Tm, ws, st, i % h->emb_slots);
        if (rc) return rc;
        FinalParams fp;
        fill_final(h, fp);
        fp.seed_ptr = seed_ptr;
        const float* z = nullptr;
        if (ddim) {
            fp.mode = 2; fp.c1 = steps[i].c1; fp.c2 = steps[i].c2; fp.c3 = steps[i].c3;
        } else {
            fp.mode = 1; fp.coef = steps[i].coef_eps; fp.div = steps[i].div; fp.sigma = steps[i].sigma;
            fp.add_noise = steps[i].add_noise ? 1 : 0;
            if (fp.add_noise) {
                if (noise_dev) z = noise_dev + (size_t)draw * n;
                fp.draw = (uint32_t)(h->noise_draw_base + draw + 1); fp.seed = seed;
                ++draw;
            }
        }
        // in-place: every thread reads only its own x element
        ScopedTimer tm(h, KC_FINAL, st);
        if (h->final_stream) FD_LAUNCH_PDL(k_final_w, dim3(L / 256, B), dim3(256), 0, st, fp, final_buffer(ws, B, Tm), (const float*)x_dev, z, x_dev,
                                      seq_dev ? seq_dev + (size_t)(i + 1) * n : (float*)nullptr, L);
        else            FD_LAUNCH_PDL(k_final, dim3(L / 256, B), dim3(256), 0, st, fp, final_buffer(ws, B, Tm), (const float*)x_dev, z, x_dev,
                                      seq_dev ? seq_dev + (size_t)(i + 1) * n : (float*)nullptr, L);
        FD_CHECK_LAUNCH(h, "k_final");
    }
    return FD_OK;
}

extern "C" int fd_sample(fd_handle* h, float* x_dev, const float* mel_dev, const fd_step* steps, int n_steps,
                         const float* noise_dev, int n_noise, uint64_t seed, int fill_xT, int ddim, float* seq_dev,
                         int B, int Tm, void* workspace_dev, size_t workspace_bytes, void* stream) {
    int rc = check_args(h, x_dev, mel_dev, steps, B, Tm, workspace_dev, workspace_bytes);
    if (rc) return rc;
    if (n_steps < 0) return fail(h, FD_ERR_INVALID, "n_steps < 0");
    int need = 0;
    if (!ddim) for (int i = 0; i < n_steps; ++i) need += steps[i].add_noise ? 1 : 0;
    if (noise_dev && n_noise < need) return fail(h, FD_ERR_INVALID, "noise_dev holds %d draws, the schedule needs %d", n_noise, need);
    FD_CUDA(h, cudaSetDevice(h->device));
    rc = setup_attrs(h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
#ifndef FD_EMU
    // ---- graph replay: device-noise mode, the whole call (N <= 64 reverse steps) captured once per (workspace, shape, schedule, options) ----
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (h->graphs && !h->timing && !noise_dev && !seq_dev && n_steps >= 1 && n_steps <= h->emb_slots && h->stop_after >= 99 &&
        h->cap && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
        const WsLayout w = ws_layout(B, Tm);
        const size_t nx = (size_t)B * Tm * HOP_TOTAL;
        fd_handle::GraphEntry* hit = nullptr;
        for (auto& g : h->gcache)
            if (g.ws == workspace_dev && g.B == B && g.Tm == Tm && g.n_steps == n_steps && g.ddim == (ddim ? 1 : 0) && g.fill_xT == (fill_xT ? 1 : 0) &&
                g.epoch == h->epoch && g.draw_base == h->noise_draw_base && !memcmp(g.steps.data(), steps, sizeof(fd_step) * n_steps)) { hit = &g; break; }
        if (!hit) {
            cudaGraph_t graph = nullptr;
            const uint64_t l0 = h->launches;
            if (cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                h->graphs = 0;   // capture unavailable in this context: plain launches from now on
                return sample_body(h, x_dev, mel_dev, steps, n_steps, noise_dev, seed, nullptr, fill_xT, ddim, seq_dev, B, Tm, ws, st);
            }
            rc = sample_body(h, ws + w.xc, ws + w.melc, steps, n_steps, nullptr, 0, h->seed_dev, fill_xT, ddim, nullptr, B, Tm, ws, h->cap);
            cudaError_t ec = cudaStreamEndCapture(h->cap, &graph);
            const uint64_t nl = h->launches - l0;
            h->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ec != cudaSuccess || !graph) return fail(h, FD_ERR_CUDA, "fd_sample: stream capture failed: %s", cudaGetErrorString(ec));
            cudaGraphExec_t exec = nullptr;
            ec = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ec != cudaSuccess) return fail(h, FD_ERR_CUDA, "fd_sample: cudaGraphInstantiate failed: %s", cudaGetErrorString(ec));
            if (h->gcache.size() >= 6) {   // drop the least recently used entry
                size_t lru = 0;
                for (size_t i = 1; i < h->gcache.size(); ++i) if (h->gcache[i].last_use < h->gcache[lru].last_use) lru = i;
                cudaGraphExecDestroy(h->gcache[lru].exec);
                h->gcache.erase(h->gcache.begin() + lru);
            }
            h->gcache.push_back({workspace_dev, B, Tm, n_steps, ddim ? 1 : 0, fill_xT ? 1 : 0, h->noise_draw_base, h->epoch, std::vector<fd_step>(steps, steps + n_steps), exec, nl, 0});
            hit = &h->gcache.back();
            ++h->graph_captures;
        }
        hit->last_use = ++h->gtick;
        FD_CUDA(h, cudaMemcpyAsync(ws + w.melc, mel_dev, (size_t)B * COND * Tm * 4, cudaMemcpyDeviceToDevice, st));
        if (!fill_xT) FD_CUDA(h, cudaMemcpyAsync(ws + w.xc, x_dev, nx * 4, cudaMemcpyDeviceToDevice, st));
        k_set_u64<<<1, 1, 0, st>>>(h->seed_dev, (unsigned long long)seed);
        FD_CUDA(h, cudaGraphLaunch(hit->exec, st));
        FD_CUDA(h, cudaMemcpyAsync(x_dev, ws + w.xc, nx * 4, cudaMemcpyDeviceToDevice, st));
        h->launches += hit->n_launches + 1;
        ++h->graph_replays;
        return FD_OK;
    }
#endif
    return sample_body(h, x_dev, mel_dev, steps, n_steps, noise_dev, seed, nullptr, fill_xT, ddim, seq_dev, B, Tm, ws, st);
}

// The step before the path: wav (B, n) -> log10-mel (B, 80, 1 + n/256), the reference's process_utterance
// (data_gen/tts/data_gen_utils.py:93-147).  fb_dev (80, 513) / range_dev (80, 2): librosa.filters.mel table and its non-zero
// bin ranges, built by the host layer (fastdiff_b200/mel.py).  Needs no weights.
extern "C" int fd_mel_frontend(fd_handle* h, const float* wav_dev, int B, int n_samples, const float* fb_dev, const int32_t* range_dev,
                               float* mel_dev, void* stream) {
    if (!h || !wav_dev || !fb_dev || !range_dev || !mel_dev || B < 1 || n_samples < 1) return fail(h, FD_ERR_INVALID, "fd_mel_frontend: bad argument");
    FD_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int Tm = 1 + n_samples / MEL_HOP;
    FD_LAUNCH(k_mel_frontend, dim3(Tm, B), dim3(256), 0, st, wav_dev, n_samples, fb_dev, (const int*)range_dev, mel_dev, Tm);
    FD_CHECK_LAUNCH(h, "k_mel_frontend");
    return FD_OK;
}

extern "C" int fd_reverse_update(fd_handle* h, float* x_dev, const float* eps_dev, const float* z_dev, const fd_step* step, int ddim,
                                 uint64_t seed, uint32_t draw, float* seq_dev, size_t count, void* stream) {
    if (!h || !x_dev || !eps_dev || !step || count < 1) return fail(h, FD_ERR_INVALID, "fd_reverse_update: bad argument");
    if (count > (size_t)0x7fffffff * 1024) return fail(h, FD_ERR_INVALID, "fd_reverse_update: count too large for one launch");
    FD_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    UpdateParams p;
    p.coef = step->coef_eps; p.div = step->div; p.sigma = step->sigma; p.c1 = step->c1; p.c2 = step->c2; p.c3 = step->c3;
    p.ddim = ddim ? 1 : 0; p.add_noise = step->add_noise ? 1 : 0; p.draw = draw; p.seed = seed;
    const uintptr_t al = (uintptr_t)x_dev | (uintptr_t)eps_dev | (uintptr_t)seq_dev;
    const unsigned blocks = (unsigned)((count + 1023) / 1024);
    FD_LAUNCH(k_reverse_update, dim3(blocks), dim3(256), 0, st, p, x_dev, eps_dev, z_dev, seq_dev, count, (al & 15) == 0 ? 1 : 0);
    FD_CHECK_LAUNCH(h, "k_reverse_update");
    return FD_OK;
}

extern "C" int fd_wav_int16(fd_handle* h, const float* x_dev, int16_t* out_dev, int B, int L, void* workspace_dev, void* stream) {
    if (!h || !x_dev || !out_dev || !workspace_dev || B < 1 || L < 1) return fail(h, FD_ERR_INVALID, "fd_wav_int16: bad argument");
    FD_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* amax = (unsigned int*)workspace_dev;   // first B words of the workspace (free between sampling calls)
    FD_CUDA(h, cudaMemsetAsync(amax, 0, (size_t)B * 4, st));
    int nb = (L + 255) / 256; if (nb > 1024) nb = 1024;
    FD_LAUNCH(k_absmax, dim3(nb, B), dim3(256), 0, st, x_dev, amax, L);
    FD_CHECK_LAUNCH(h, "k_absmax");
    FD_LAUNCH(k_wav_int16, dim3((L + 255) / 256, B), dim3(256), 0, st, x_dev, (const unsigned int*)amax, out_dev, L);
    FD_CHECK_LAUNCH(h, "k_wav_int16");
    return FD_OK;
}

extern "C" int fd_debug_read(fd_handle* h, const char* name, float* out_dev, size_t* count, int B, int Tm,
                             void* workspace_dev, void* stream) {
    if (!h || !name || !count || !workspace_dev) return fail(h, FD_ERR_INVALID, "fd_debug_read: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace_dev;
    const WsLayout w = ws_layout(B, Tm);
    const size_t L = (size_t)Tm * HOP_TOTAL;
    const size_t len = strlen(name);
    const int n = (len > 0 && name[len - 1] >= '0' && name[len - 1] <= '2') ? name[len - 1] - '0' : -1;
    auto gather = [&](const float* src, int T, int Cn, int row_stride, size_t item_stride, int off) -> int {
        *count = (size_t)B * Cn * T;
        if (!out_dev) return FD_OK;
        FD_LAUNCH(k_cl_to_ncl, dim3((unsigned)((*count + 255) / 256)), dim3(256), 0, st, src, out_dev, B, T, Cn, row_stride, (int)item_stride, off);
        FD_CHECK_LAUNCH(h, "k_cl_to_ncl");
        return FD_OK;
    };
#ifndef FD_EMU
    if (!strcmp(name, "lvc_timeline")) {   // 128 clock64 stamps as (hi,lo)-free doubles are overkill: deltas fit fp32
        *count = 128;
        if (!out_dev) return FD_OK;
        unsigned long long host[128];
        FD_CUDA(h, cudaStreamSynchronize(st));
        FD_CUDA(h, cudaMemcpyFromSymbol(host, g_lvc_timeline, sizeof host));
        float rel[128];
        for (int i = 0; i < 128; ++i) rel[i] = (float)(double)(host[i] - host[0]);
        FD_CUDA(h, cudaMemcpyAsync(out_dev, rel, sizeof rel, cudaMemcpyHostToDevice, st));
        FD_CUDA(h, cudaStreamSynchronize(st));
        return FD_OK;
    }
#if defined(B0_TIMELINE)
    if (!strcmp(name, "b0_timeline")) {   // k_lvc_layer_b0h phase timeline
        constexpr int NW = 4 * 8;
        *count = NW;
        if (!out_dev) return FD_OK;
        static unsigned long long host[NW];
        static float rel[NW];
        FD_CUDA(h, cudaStreamSynchronize(st));
        FD_CUDA(h, cudaMemcpyFromSymbol(host, g_b0_timeline, sizeof host));
        unsigned long long mn = ~0ull;
        for (int i = 0; i < NW; ++i) if (host[i] && host[i] < mn) mn = host[i];
        for (int i = 0; i < NW; ++i) rel[i] = host[i] ? (float)(double)(host[i] - mn + 1) : 0.f;
        FD_CUDA(h, cudaMemcpyAsync(out_dev, rel, sizeof rel, cudaMemcpyHostToDevice, st));
        FD_CUDA(h, cudaStreamSynchronize(st));
        return FD_OK;
    }
#endif
#if defined(UT_TIMELINE)
    if (!strcmp(name, "ut_timeline")) {   // k_upsample_tc<4, POUT> phase timeline
        constexpr int NW = 24 * 8;
        *count = NW;
        if (!out_dev) return FD_OK;
        static unsigned long long host[NW];
        static float rel[NW];
        FD_CUDA(h, cudaStreamSynchronize(st));
        FD_CUDA(h, cudaMemcpyFromSymbol(host, g_ut_timeline, sizeof host));
        unsigned long long mn = ~0ull;
        for (int i = 0; i < NW; ++i) if (host[i] && host[i] < mn) mn = host[i];
        for (int i = 0; i < NW; ++i) rel[i] = host[i] ? (float)(double)(host[i] - mn + 1) : 0.f;
        FD_CUDA(h, cudaMemcpyAsync(out_dev, rel, sizeof rel, cudaMemcpyHostToDevice, st));
        FD_CUDA(h, cudaStreamSynchronize(st));
        return FD_OK;
    }
#endif
#if defined(KC_TIMELINE)
    if (!strcmp(name, "kc_timeline")) {   // kernel_conv GEMM role timeline (fd_kernels_tc.cuh): cycles relative to the earliest stamp, 0 where unset
        constexpr int NW = 3 * KC_TL_ITEMS * 8;
        *count = NW;
        if (!out_dev) return FD_OK;
        static unsigned long long host[NW];
        static float rel[NW];
        FD_CUDA(h, cudaStreamSynchronize(st));
        FD_CUDA(h, cudaMemcpyFromSymbol(host, g_kc_timeline, sizeof host));
        unsigned long long mn = ~0ull;
        for (int i = 0; i < NW; ++i) if (host[i] && host[i] < mn) mn = host[i];
        for (int i = 0; i < NW; ++i) rel[i] = host[i] ? (float)(double)(host[i] - mn + 1) : 0.f;
        FD_CUDA(h, cudaMemcpyAsync(out_dev, rel, sizeof rel, cudaMemcpyHostToDevice, st));
        FD_CUDA(h, cudaStreamSynchronize(st));
        return FD_OK;
    }
#endif
#if defined(LP_TIMELINE)
    if (!strcmp(name, "lp_timeline")) {   // k_lvc_p role timeline (fd_kernels_lvcp.cuh): cycles relative to the earliest stamp, 0 where unset
        constexpr int NW = 4 * LP_TL_TILES * 8;
        *count = NW;
        if (!out_dev) return FD_OK;
        static unsigned long long host[NW];
        static float rel[NW];
        FD_CUDA(h, cudaStreamSynchronize(st));
        FD_CUDA(h, cudaMemcpyFromSymbol(host, g_lp_timeline, sizeof host));
        unsigned long long mn = ~0ull;
        for (int i = 0; i < NW; ++i) if (host[i] && host[i] < mn) mn = host[i];
        for (int i = 0; i < NW; ++i) rel[i] = host[i] ? (float)(double)(host[i] - mn + 1) : 0.f;
        FD_CUDA(h, cudaMemcpyAsync(out_dev, rel, sizeof rel, cudaMemcpyHostToDevice, st));
        FD_CUDA(h, cudaStreamSynchronize(st));
        return FD_OK;
    }
#endif
#endif
    if (!strcmp(name, "embed")) {
        *count = (size_t)B * EMB_OUT;
        if (out_dev) FD_CUDA(h, cudaMemcpyAsync(out_dev, ws + w.emb, *count * 4, cudaMemcpyDeviceToDevice, st));
        return FD_OK;
    }
    if (!strncmp(name, "down", 4) && n >= 0) {
        const size_t T = n == 0 ? L / 4 : (n == 1 ? L / 32 : L / 256);
        const float* src = ws + (n == 0 ? w.d0 : (n == 1 ? w.d1 : w.d2));
        return gather(src, (int)T, C, C, T * C, 0);
    }
    if (!strncmp(name, "kp_hidden", 9) && n >= 0)
        return gather(ws + w.hk + (size_t)n * B * (Tm + 2) * HID, Tm, HID, HID, (size_t)(Tm + 2) * HID, HID);
    if (!strncmp(name, "lvc", 3) && n >= 0) {
        const size_t T = (size_t)Tm * hop_of(n);
        return gather(block_out_buffer(ws, B, Tm, n), (int)T, C, C, T * C, 0);
    }
    if ((!strncmp(name, "kernels", 7) || !strncmp(name, "kbias", 5)) && n >= 0) {
        const int want_bias = name[1] == 'b';
        *count = want_bias ? (size_t)B * LAYERS * LVC_OUT * Tm : (size_t)B * LAYERS * C * LVC_OUT * KS * Tm;
        if (!out_dev) return FD_OK;
        FD_LAUNCH(k_kern_to_ref, dim3((unsigned)((*count + 255) / 256)), dim3(256), 0, st, ws + w.kern + (size_t)n * B * Tm * KCN, out_dev, B, Tm, want_bias,
                  n == 0 ? (h->b0_converted > 0 ? 2 : 1) : (h->mode == FD_MODE_TC_3XF16 ? (h->kimg_last ? 3 : 2) : 0));
        FD_CHECK_LAUNCH(h, "k_kern_to_ref");
        return FD_OK;
    }
    return fail(h, FD_ERR_INVALID, "fd_debug_read: unknown tensor '%s'", name);
}
