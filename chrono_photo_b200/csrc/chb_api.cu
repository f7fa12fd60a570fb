// chb_api.cu -- C ABI of libchrono_b200.so (see include/chrono_b200.h). Host side: contexts, the HBM-resident stack,
// the asynchronous ingest pipeline and kernel dispatch. No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/chrono_b200.h"
#include "chb_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <condition_variable>
#include <map>
#include <tuple>
#include <vector>

using namespace chb;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static thread_local uint64_t g_last_slow = 0, g_last_hard = 0;
static thread_local float g_last_main_ms = 0.0f;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) return fail(CHB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---- tuning / test knobs: read from the environment ONCE when the library is loaded (no getenv on the launch path); tests and
// tuning scripts change them through chb_set_tuning. -1 = automatic.
struct Tuning {
    std::atomic<int> force_variant{-1};    // CHB_FORCE_VARIANT: K1 variant table index (a larger-capacity variant than needed)
    std::atomic<int> hist{-1};             // CHB_HIST: 0 / 1 forces the solver / the histogram kernel on the iterative tier
    std::atomic<int> pdl{1};               // CHB_PDL: tier kernels as programmatic dependent launches
    std::atomic<int> video_queue_cap{-1};  // CHB_VIDEO_QUEUE_CAP: capacity of the video exact-path queue (tests the in-place fallback)
    std::atomic<int> inline_min{12};       // CHB_INLINE_MIN: uncertified pixels per tile from which the tile is finished inside K1 (0 = never)
    std::atomic<int> hard_inline_min{1};   // CHB_HARD_INLINE_MIN: same for a warp-full of the iterative tier (finished inside outlier_hard_kernel)
    std::atomic<int> hard_window{1};       // CHB_HARD_WINDOW: the iterative tier tries the two straight-line windows before the solver
    std::atomic<int> hard_drains{1};       // CHB_HARD_DRAINS: outlier_hard_kernel also drains the streaming kernel's own queue (no exact-path launch)
    Tuning() {
        auto env = [](const char* k, std::atomic<int>& v) { if (const char* e = getenv(k)) v.store(atoi(e)); };
        env("CHB_FORCE_VARIANT", force_variant); env("CHB_HIST", hist); env("CHB_PDL", pdl);
        env("CHB_VIDEO_QUEUE_CAP", video_queue_cap); env("CHB_INLINE_MIN", inline_min); env("CHB_HARD_INLINE_MIN", hard_inline_min); env("CHB_HARD_DRAINS", hard_drains); env("CHB_HARD_WINDOW", hard_window);
    }
};
static Tuning g_tune;
extern "C" int chb_set_tuning(const char* key, int value) {
    if (!key) return fail(CHB_ERR_INVALID, "chb_set_tuning: null key");
    const std::string k(key);
    if (k == "force_variant") g_tune.force_variant.store(value);
    else if (k == "hist") g_tune.hist.store(value);
    else if (k == "pdl") g_tune.pdl.store(value);
    else if (k == "video_queue_cap") g_tune.video_queue_cap.store(value);
    else if (k == "inline_min") g_tune.inline_min.store(value);
    else if (k == "hard_inline_min") g_tune.hard_inline_min.store(value);
    else if (k == "hard_drains") g_tune.hard_drains.store(value);
    else if (k == "hard_window") g_tune.hard_window.store(value);
    else return fail(CHB_ERR_INVALID, "chb_set_tuning: unknown key '%s'", key);
    return CHB_OK;
}

extern "C" const char* chb_last_error(void) { return g_err.c_str(); }
extern "C" int chb_version(void) { return CHB_VERSION; }
extern "C" uint64_t chb_launch_count(void) { return g_launches.load(); }
extern "C" void chb_launch_count_reset(void) { g_launches.store(0); }
extern "C" uint64_t chb_last_slow_pixels(void) { return g_last_slow; }
extern "C" uint64_t chb_last_hard_pixels(void) { return g_last_hard; }
extern "C" float chb_last_main_kernel_ms(void) { return g_last_main_ms; }

// ------------------------------------------------------------------------------------------------ objects
struct Dev {
    int id = 0;
    int sm_count = 148;
    cudaStream_t compute = nullptr, own_compute = nullptr, copy = nullptr, pack = nullptr;
};
struct chb_ctx {
    std::vector<Dev> devs;
};

// Per-band device state of ONE compositing call in flight. A stack owns kCallSlots of them, so that chb_outlier / chb_simple
// entered from several host threads (the reference calls process() from a rayon pool: src/main.rs:260-261, :378-379) overlap their
// launches, tier kernels and D2H copies instead of queueing on one output buffer. Allocated at first use of the slot.
static constexpr int kCallSlots = 4;
struct CallBuf {
    cudaStream_t own_stream = nullptr;  // slots >= 1; slot 0 runs on the device's compute stream (chb_ctx_set_stream)
    uint8_t* d_out = nullptr;
    uint8_t* d_mask = nullptr;
    uint32_t* d_wmask = nullptr;
    uint32_t* d_smask = nullptr;
    int32_t* d_win = nullptr;
    int32_t* d_posg = nullptr;
    float* d_fade = nullptr;
    unsigned long long* d_counters = nullptr;
    float *d_dbg_median = nullptr, *d_dbg_q1 = nullptr, *d_dbg_q3 = nullptr;
    int* d_dbg_nout = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;  // ev_mid: after the streaming kernel, before the tier kernels
    QueueEntry* d_queue = nullptr;        // exact-path queue of an outlier launch + its counters ([0..1] exact, [2..3] iterative tier)
    long long* d_hqueue = nullptr;        // iterative-tier queue (pixel indices)
    uint32_t* d_hflags = nullptr;         // per-tile flag words the iterative-tier queue is built from
    unsigned int* d_qcount = nullptr;     // two sets of 16 words (queue counters + the call's four 64-bit counters), then the flag words
    int qset = 0;                         // the set the NEXT call uses; it is all zero (the previous call's compaction kernel cleared it)
    int last_qset = 0;                    // the set the last call used (its counters are fetched from there)
};

struct Band {
    int dev_slot = 0;
    int row0 = 0, rows = 0;
    long long n_pixels = 0, n_tiles = 0;
    size_t stack_bytes = 0, frame_bytes = 0;
    uint8_t* d_stack = nullptr;
    CallBuf call[kCallSlots];
    // ingest: device staging for two frame groups (2 x 16 frames, allocated at the first upload). A group whose frames have all
    // arrived is re-laid-out by ONE pack_group_kernel launch that writes whole 16-byte units (no byte scatter); the other
    // region receives the next group's copies meanwhile.
    uint8_t* d_stage = nullptr;  // [2 * 16][frame_bytes]
    struct Region { int group = -1; uint32_t arrived = 0; } region[2];
    cudaEvent_t copied = nullptr;               // recorded on d.copy after a group's last H2D (stream order covers the earlier ones)
    cudaEvent_t packed[2] = {nullptr, nullptr};  // the pack launch that last read region r finished
    uint8_t* d_tmp = nullptr;                    // one frame: chb_stack_download
    // chrono-video (runs on call slot 0): two slots of window planes (composites, masks), per-window warning counters
    uint8_t* d_vout[2] = {nullptr, nullptr};
    uint8_t* d_vmask[2] = {nullptr, nullptr};
    int v_cap_windows = 0;
    unsigned long long* d_vwarn = nullptr;
    int vwarn_cap = 0;
    VideoQueueEntry* d_vqueue = nullptr;  // exact-path queue of a chunk + its counter
    unsigned int* d_vqcount = nullptr;
    cudaEvent_t v_done[2] = {nullptr, nullptr};    // kernel writing slot s finished
    cudaEvent_t v_copied[2] = {nullptr, nullptr};  // D2H of slot s finished
};

// Host side of a call slot: pinned scratch for the per-call tables and what the slot's last call left behind.
struct CallSlot {
    bool ready = false, in_use = false;
    uint32_t* h_wmask = nullptr;
    uint32_t* h_smask = nullptr;
    int32_t* h_win = nullptr;
    int32_t* h_posg = nullptr;
    float* h_fade = nullptr;
    unsigned long long* h_counters = nullptr;  // [n_bands * 4]
    int last_kind = 0;  // what ran last on the slot: 0 nothing yet, 1 a single-window call (its planes can be fetched), 2 a video run
    bool last_has_mask = false;
    bool counters_pending = false;  // the last enqueued call left its counters on the devices: collect_outlier fetches them
    uint64_t last_warnings = 0;
    std::vector<uint8_t> last_tables;  // fingerprint of the per-call tables currently on the devices
};

struct chb_stack {
    chb_ctx* ctx = nullptr;
    int W = 0, H = 0, C = 0, N = 0, NG = 0;
    std::vector<Band> bands;
    std::vector<uint8_t> uploaded;  // per frame
    std::mutex upload_mu;
    // call slots: the host-buffer entry points (chb_outlier, chb_simple) take any free slot; the device-side API (chb_*_device,
    // chb_outlier_enqueue, chb_stack_wait, chb_fetch_last*, chb_outlier_video*) is defined on "the last call" and owns slot 0
    CallSlot slots[kCallSlots];
    std::mutex call_mu;
    std::condition_variable call_cv;
    // pageable sources: a pool of pinned host frames; a caller thread owns one slot while it copies its frame into it (no lock
    // held during the memcpy), so several decode threads stage in parallel
    static constexpr int kHostSlots = 8;
    uint8_t* h_slot[kHostSlots] = {};
    bool h_in_use[kHostSlots] = {};
    std::vector<cudaEvent_t> h_done[kHostSlots];  // per band: the H2D out of the slot finished
    bool h_pending[kHostSlots] = {};
    std::mutex slot_mu;
    std::condition_variable slot_cv;
};

static constexpr int kMaxWindowFrames = 4096;  // largest window span the register-resident kernels hold (longer whole-stack series: histogram tier)
static constexpr int kMaxGroupsTable = kMaxWindowFrames / 16;

// ------------------------------------------------------------------------------------------------ context
extern "C" int chb_ctx_create(const int* device_ids, int n_dev, chb_ctx** out) {
    if (!out) return fail(CHB_ERR_INVALID, "chb_ctx_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(CHB_ERR_CUDA, "chb_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    std::vector<int> ids;
    if (!device_ids || n_dev <= 0) ids.push_back(0);
    else ids.assign(device_ids, device_ids + n_dev);
    chb_ctx* ctx = new chb_ctx();
    for (int id : ids) {
        if (id < 0 || id >= count) {
            delete ctx;
            return fail(CHB_ERR_INVALID, "chb_ctx_create: device %d out of range (have %d)", id, count);
        }
        Dev d;
        d.id = id;
        CU(cudaSetDevice(id));
        cudaDeviceProp pr;
        CU(cudaGetDeviceProperties(&pr, id));
        if (pr.major < 10) {
            delete ctx;
            return fail(CHB_ERR_UNSUPPORTED, "chb_ctx_create: device %d is sm_%d%d; this build carries sm_100a code only", id, pr.major, pr.minor);
        }
        d.sm_count = pr.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&d.own_compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&d.pack, cudaStreamNonBlocking));
        d.compute = d.own_compute;
        ctx->devs.push_back(d);
    }
    *out = ctx;
    return CHB_OK;
}

extern "C" int chb_ctx_destroy(chb_ctx* ctx) {
    if (!ctx) return CHB_OK;
    for (Dev& d : ctx->devs) {
        cudaSetDevice(d.id);
        cudaStreamDestroy(d.own_compute);
        cudaStreamDestroy(d.copy);
        cudaStreamDestroy(d.pack);
    }
    delete ctx;
    return CHB_OK;
}

extern "C" int chb_ctx_device_count(const chb_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

extern "C" int chb_ctx_mem_info(chb_ctx* ctx, int dev_slot, size_t* free_bytes, size_t* total_bytes) {
    if (!ctx || dev_slot < 0 || dev_slot >= (int)ctx->devs.size()) return fail(CHB_ERR_INVALID, "chb_ctx_mem_info: bad device slot");
    size_t f = 0, t = 0;
    CU(cudaSetDevice(ctx->devs[dev_slot].id));
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return CHB_OK;
}

extern "C" int chb_ctx_set_stream(chb_ctx* ctx, int dev_slot, void* cuda_stream) {
    if (!ctx || dev_slot < 0 || dev_slot >= (int)ctx->devs.size()) return fail(CHB_ERR_INVALID, "chb_ctx_set_stream: bad device slot");
    Dev& d = ctx->devs[dev_slot];
    d.compute = cuda_stream ? (cudaStream_t)cuda_stream : d.own_compute;
    return CHB_OK;
}

// ------------------------------------------------------------------------------------------------ stack
static void free_band(Band& b) {
    cudaFree(b.d_stack);
    cudaFree(b.d_stage); cudaFree(b.d_tmp);
    if (b.copied) cudaEventDestroy(b.copied);
    for (int s = 0; s < 2; s++)
        if (b.packed[s]) cudaEventDestroy(b.packed[s]);
    for (CallBuf& cb : b.call) {
        cudaFree(cb.d_out); cudaFree(cb.d_mask);
        cudaFree(cb.d_wmask); cudaFree(cb.d_smask); cudaFree(cb.d_win); cudaFree(cb.d_posg); cudaFree(cb.d_fade); cudaFree(cb.d_counters);
        cudaFree(cb.d_dbg_median); cudaFree(cb.d_dbg_q1); cudaFree(cb.d_dbg_q3); cudaFree(cb.d_dbg_nout);
        cudaFree(cb.d_queue); cudaFree(cb.d_hqueue); cudaFree(cb.d_qcount);
        if (cb.ev0) cudaEventDestroy(cb.ev0);
        if (cb.ev1) cudaEventDestroy(cb.ev1);
        if (cb.ev_mid) cudaEventDestroy(cb.ev_mid);
        if (cb.own_stream) cudaStreamDestroy(cb.own_stream);
    }
    for (int s = 0; s < 2; s++) {
        cudaFree(b.d_vout[s]); cudaFree(b.d_vmask[s]);
        if (b.v_done[s]) cudaEventDestroy(b.v_done[s]);
        if (b.v_copied[s]) cudaEventDestroy(b.v_copied[s]);
    }
    cudaFree(b.d_vwarn); cudaFree(b.d_vqueue); cudaFree(b.d_vqcount);
}

extern "C" int chb_stack_destroy(chb_stack* st) {
    if (!st) return CHB_OK;
    for (Band& b : st->bands) {
        cudaSetDevice(st->ctx->devs[b.dev_slot].id);
        cudaDeviceSynchronize();
        free_band(b);
    }
    for (int k = 0; k < chb_stack::kHostSlots; k++) {
        if (st->h_slot[k]) cudaFreeHost(st->h_slot[k]);
        for (cudaEvent_t e : st->h_done[k]) cudaEventDestroy(e);
    }
    for (CallSlot& cs : st->slots) {
        if (cs.h_wmask) cudaFreeHost(cs.h_wmask);
        if (cs.h_smask) cudaFreeHost(cs.h_smask);
        if (cs.h_win) cudaFreeHost(cs.h_win);
        if (cs.h_posg) cudaFreeHost(cs.h_posg);
        if (cs.h_fade) cudaFreeHost(cs.h_fade);
        if (cs.h_counters) cudaFreeHost(cs.h_counters);
    }
    delete st;
    return CHB_OK;
}

// Buffers of call slot `slot` (pinned tables, per band: output planes, tables, events, a stream for slots >= 1), at first use.
// The caller owns the slot.
static int ensure_call_slot(chb_stack* st, int slot) {
    CallSlot& cs = st->slots[slot];
    if (cs.ready) return CHB_OK;
    const int nd = (int)st->bands.size();
    CU(cudaMallocHost(&cs.h_wmask, sizeof(uint32_t) * kMaxGroupsTable * 4));
    CU(cudaMallocHost(&cs.h_smask, sizeof(uint32_t) * kMaxGroupsTable * 4));
    CU(cudaMallocHost(&cs.h_win, sizeof(int32_t) * (size_t)std::max(st->N, 1)));
    CU(cudaMallocHost(&cs.h_posg, sizeof(int32_t) * (size_t)st->NG));
    CU(cudaMallocHost(&cs.h_fade, sizeof(float) * CHB_MAX_FADE_VALUES));
    CU(cudaMallocHost(&cs.h_counters, sizeof(unsigned long long) * 4 * nd));
    for (Band& b : st->bands) {
        CallBuf& cb = b.call[slot];
        CU(cudaSetDevice(st->ctx->devs[b.dev_slot].id));
        if (slot > 0) CU(cudaStreamCreateWithFlags(&cb.own_stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&cb.d_out, b.frame_bytes));
        CU(cudaMalloc(&cb.d_mask, b.frame_bytes));
        CU(cudaMalloc(&cb.d_wmask, sizeof(uint32_t) * kMaxGroupsTable * 4));
        CU(cudaMalloc(&cb.d_smask, sizeof(uint32_t) * kMaxGroupsTable * 4));
        CU(cudaMalloc(&cb.d_win, sizeof(int32_t) * (size_t)st->N));
        CU(cudaMalloc(&cb.d_posg, sizeof(int32_t) * (size_t)st->NG));
        CU(cudaMalloc(&cb.d_fade, sizeof(float) * CHB_MAX_FADE_VALUES));
        CU(cudaMalloc(&cb.d_counters, sizeof(unsigned long long) * 4));
        CU(cudaEventCreate(&cb.ev0));
        CU(cudaEventCreate(&cb.ev1));
        CU(cudaEventCreate(&cb.ev_mid));
    }
    cs.ready = true;
    return CHB_OK;
}

// Takes a call slot for the duration of one entry point: slot 0 (the device-side API's "last call") or the lowest free one.
struct SlotGuard {
    chb_stack* st;
    int slot = -1;
    SlotGuard(chb_stack* s, bool want_zero) : st(s) {
        std::unique_lock<std::mutex> lk(st->call_mu);
        if (want_zero) {
            st->call_cv.wait(lk, [&] { return !st->slots[0].in_use; });
            slot = 0;
        } else {
            st->call_cv.wait(lk, [&] { for (const CallSlot& c : st->slots) if (!c.in_use) return true; return false; });
            slot = 0;
            while (st->slots[slot].in_use) slot++;
        }
        st->slots[slot].in_use = true;
    }
    ~SlotGuard() {
        { std::lock_guard<std::mutex> lk(st->call_mu); st->slots[slot].in_use = false; }
        st->call_cv.notify_all();
    }
};
static cudaStream_t call_stream(const Dev& d, const CallBuf& cb, int slot) { return slot == 0 ? d.compute : cb.own_stream; }

extern "C" int chb_stack_create(chb_ctx* ctx, int width, int height, int channels, int n_frames, chb_stack** out) {
    if (!ctx || !out) return fail(CHB_ERR_INVALID, "chb_stack_create: null argument");
    if (width < 1 || height < 1 || n_frames < 1) return fail(CHB_ERR_INVALID, "chb_stack_create: width, height and n_frames must be positive");
    if (channels != 3 && channels != 4)
        return fail(CHB_ERR_UNSUPPORTED, "chb_stack_create: %d channels; the reference only writes Rgb8/Rgba8 (src/main.rs:550-567)", channels);
    chb_stack* st = new chb_stack();
    st->ctx = ctx;
    st->W = width; st->H = height; st->C = channels; st->N = n_frames;
    st->NG = (n_frames + kGroupFrames - 1) / kGroupFrames;
    st->uploaded.assign(n_frames, 0);
    const int nd = std::min<int>((int)ctx->devs.size(), height);
    auto bail = [&](int code) { chb_stack_destroy(st); return code; };
#define CUB(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess) return bail(fail(CHB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__)); \
    } while (0)
    st->bands.resize(nd);
    for (int k = 0; k < nd; k++) {
        Band& b = st->bands[k];
        b.dev_slot = k;
        b.row0 = (int)((long long)height * k / nd);  // row shard: GPU k owns rows [H*k/G, H*(k+1)/G)
        b.rows = (int)((long long)height * (k + 1) / nd) - b.row0;
        b.n_pixels = (long long)b.rows * width;
        b.n_tiles = (b.n_pixels + kTilePixels - 1) / kTilePixels;
        b.stack_bytes = (size_t)(b.n_tiles * tile_bytes(channels, st->NG));
        b.frame_bytes = (size_t)b.n_pixels * channels;
        CUB(cudaSetDevice(ctx->devs[k].id));
        CUB(cudaMalloc(&b.d_stack, b.stack_bytes));
        CUB(cudaMemset(b.d_stack, 0, b.stack_bytes));  // frames beyond n_frames in the last group must read as zero
        CUB(cudaEventCreateWithFlags(&b.copied, cudaEventDisableTiming));
        for (int s = 0; s < 2; s++) CUB(cudaEventCreateWithFlags(&b.packed[s], cudaEventDisableTiming));
        CUB(cudaDeviceSynchronize());
    }
#undef CUB
    {
        const int rc = ensure_call_slot(st, 0);
        if (rc) return bail(rc);
    }
    *out = st;
    return CHB_OK;
}

extern "C" size_t chb_stack_device_bytes(const chb_stack* st, int dev_slot) {
    if (!st || dev_slot < 0 || dev_slot >= (int)st->bands.size()) return 0;
    return st->bands[dev_slot].stack_bytes;
}

// Launch attributes and occupancy of a (kernel, device, dynamic smem) triple are set / queried once and remembered: the launch
// path of a compositing call makes no cudaFuncSetAttribute / occupancy call after the first use (host time that shows at the
// size of config 1 and of an eighth of config 3).
static int kernel_config(const void* kern, int device, int threads, int smem, int* occ_out) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, int>, int> cache;
    static std::map<std::pair<const void*, int>, int> max_smem;  // the attribute is per function: it is only ever raised (a
                                                                 // kernel launched with two sizes -- the chrono-video kernel with one
                                                                 // or two result words -- must keep the larger limit)
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(kern, device, smem);
    auto it = cache.find(key);
    if (it == cache.end()) {
        int occ = 1;
        int& cur = max_smem[std::make_pair(kern, device)];
        if (smem > cur || cur == 0) {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, cur)));
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            cur = std::max(smem, cur);
        }
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        it = cache.emplace(key, std::max(1, occ)).first;
    }
    if (occ_out) *occ_out = it->second;
    return CHB_OK;
}

static int grid_for(long long work_items, int threads, int sm_count, int waves) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)sm_count * waves;
    return (int)std::max<long long>(1, std::min(blocks, cap));
}

// ---- ingest (replaces TimeSlicer::write_time_slices, src/slicer.rs:106-231, and the chunk streams, src/streams.rs:93-202) ----
constexpr int kStageFrames = 2 * kGroupFrames;

// Re-layout of whatever has arrived in region r. complete: every frame of the group is there -> one launch writing whole units
// (frames beyond n_frames in the last group are written as zeros); otherwise the present frames are scattered one by one
// (read-modify-write of their byte inside each unit: slow, but only taken by out-of-order or partial uploads).
static int pack_region(chb_stack* st, Band& b, Dev& d, int r) {
    Band::Region& R = b.region[r];
    if (R.group < 0 || R.arrived == 0) { R.group = -1; R.arrived = 0; return CHB_OK; }
    const int g = R.group;
    const int in_group = std::min(kGroupFrames, st->N - g * kGroupFrames);
    const uint32_t full = in_group >= 32 ? 0xffffffffu : ((1u << in_group) - 1u);
    CU(cudaEventRecord(b.copied, d.copy));
    CU(cudaStreamWaitEvent(d.pack, b.copied, 0));
    if (R.arrived == full) {
        PackGroupArgs pa;
        for (int k = 0; k < kGroupFrames; k++) pa.src[k] = k < in_group ? b.d_stage + (size_t)(r * kGroupFrames + k) * b.frame_bytes : nullptr;
        pack_group_kernel<<<grid_for(b.n_pixels, 256, d.sm_count, 8), 256, 0, d.pack>>>(pa, b.d_stack, b.n_pixels, st->C, st->NG, g);
        g_launches++;
    } else {
        for (int k = 0; k < kGroupFrames; k++) {
            if (!((R.arrived >> k) & 1u)) continue;
            pack_frame_kernel<<<grid_for(b.n_pixels, 256, d.sm_count, 8), 256, 0, d.pack>>>(b.d_stage + (size_t)(r * kGroupFrames + k) * b.frame_bytes, b.d_stack,
                                                                                          b.n_pixels, st->C, st->NG, g * kGroupFrames + k);
            g_launches++;
        }
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(b.packed[r], d.pack));
    R.group = -1;
    R.arrived = 0;
    return CHB_OK;
}

// Enqueues the re-layout of every frame that has been uploaded but not packed yet (incomplete groups). The caller holds upload_mu.
static int flush_ingest_locked(chb_stack* st) {
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        if (b.region[0].arrived == 0 && b.region[1].arrived == 0) continue;
        CU(cudaSetDevice(d.id));
        for (int r = 0; r < 2; r++) {
            int rc = pack_region(st, b, d, r);
            if (rc) return rc;
        }
    }
    return CHB_OK;
}

// Orders a compute stream after every ingest step enqueued so far on this band (H2D copies on d.copy, re-layout on d.pack), so a
// compositing launch issued right after a run of uploads never reads a partly packed stack -- with or without chb_stack_sync.
// (Waiting on an event that was never recorded is a no-op.) Call flush_ingest first.
static cudaError_t wait_ingest(Band& b, cudaStream_t s) {
    for (int k = 0; k < 2; k++) {
        if (!b.packed[k]) continue;
        cudaError_t e = cudaStreamWaitEvent(s, b.packed[k], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
static int flush_ingest(chb_stack* st) {
    std::lock_guard<std::mutex> lk(st->upload_mu);
    return flush_ingest_locked(st);
}

// A pinned host frame for a pageable source: blocks until a slot of the pool is free and its previous DMA has left.
static int acquire_host_slot(chb_stack* st, int& slot) {
    {
        std::unique_lock<std::mutex> lk(st->slot_mu);
        st->slot_cv.wait(lk, [&] { for (bool u : st->h_in_use) if (!u) return true; return false; });
        slot = 0;
        while (st->h_in_use[slot]) slot++;
        st->h_in_use[slot] = true;
    }
    auto release = [&](int code) {
        std::lock_guard<std::mutex> lk(st->slot_mu);
        st->h_in_use[slot] = false;
        st->slot_cv.notify_one();
        return code;
    };
    const size_t bytes = (size_t)st->W * st->H * st->C;
    if (!st->h_slot[slot]) {
        cudaError_t e = cudaMallocHost(&st->h_slot[slot], bytes);
        if (e != cudaSuccess) return release(fail(CHB_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)));
        st->h_done[slot].resize(st->bands.size(), nullptr);
        for (size_t k = 0; k < st->bands.size(); k++) {
            cudaSetDevice(st->ctx->devs[st->bands[k].dev_slot].id);
            e = cudaEventCreateWithFlags(&st->h_done[slot][k], cudaEventDisableTiming);
            if (e != cudaSuccess) return release(fail(CHB_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)));
        }
    }
    if (st->h_pending[slot]) {
        for (cudaEvent_t ev : st->h_done[slot]) {
            cudaError_t e = cudaEventSynchronize(ev);
            if (e != cudaSuccess) return release(fail(CHB_ERR_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(e)));
        }
        st->h_pending[slot] = false;
    }
    return CHB_OK;
}
static void release_host_slot(chb_stack* st, int slot) {
    std::lock_guard<std::mutex> lk(st->slot_mu);
    st->h_in_use[slot] = false;
    st->slot_cv.notify_one();
}

// The bookkeeping every ingest path shares: picks the frame's staging slot in its group's region, lets `enqueue` put the band's
// rows there (a copy enqueued on d.copy), and launches the group's re-layout once its last frame has arrived.
template <typename Enqueue>
static int stage_frame(chb_stack* st, int frame_idx, Enqueue&& enqueue) {
    int rc = CHB_OK;
    std::lock_guard<std::mutex> lk(st->upload_mu);
    const int g = frame_idx / kGroupFrames, k = frame_idx % kGroupFrames, r = g & 1;
    for (size_t bi = 0; bi < st->bands.size() && rc == CHB_OK; bi++) {
        Band& b = st->bands[bi];
        Dev& d = st->ctx->devs[b.dev_slot];
        auto cu = [&](cudaError_t e, const char* what) {
            if (e != cudaSuccess && rc == CHB_OK) rc = fail(CHB_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
            return e == cudaSuccess;
        };
        if (!cu(cudaSetDevice(d.id), "cudaSetDevice")) break;
        if (!b.d_stage && !cu(cudaMalloc(&b.d_stage, (size_t)kStageFrames * b.frame_bytes), "cudaMalloc(ingest staging)")) break;
        Band::Region& R = b.region[r];
        if ((R.group >= 0 && R.group != g) || ((R.arrived >> k) & 1u)) {  // the region holds another group, or this frame again
            rc = pack_region(st, b, d, r);
            if (rc) break;
        }
        R.group = g;
        // the pack launch that last read this region must be done before a copy overwrites it
        if (!cu(cudaStreamWaitEvent(d.copy, b.packed[r], 0), "cudaStreamWaitEvent")) break;
        if (!cu(enqueue(bi, b, d, b.d_stage + (size_t)(r * kGroupFrames + k) * b.frame_bytes), "ingest copy")) break;
        R.arrived |= 1u << k;
        const int in_group = std::min(kGroupFrames, st->N - g * kGroupFrames);
        if (R.arrived == ((1u << in_group) - 1u)) rc = pack_region(st, b, d, r);
    }
    if (rc == CHB_OK) st->uploaded[frame_idx] = 1;
    return rc;
}

// Ingest of one frame into every band. `pinned`: the source can be DMA'd directly; otherwise it is first copied into a pinned
// slot of the pool by the calling thread (in parallel with other callers).
static int upload_impl(chb_stack* st, int frame_idx, const uint8_t* host, size_t pitch, int crop_x, int crop_y, bool pinned) {
    if (!st || !host) return fail(CHB_ERR_INVALID, "chb_stack_upload: null argument");
    if (frame_idx < 0 || frame_idx >= st->N) return fail(CHB_ERR_INVALID, "chb_stack_upload: frame %d outside [0, %d)", frame_idx, st->N);
    if (crop_x < 0 || crop_y < 0) return fail(CHB_ERR_INVALID, "chb_stack_upload: negative crop origin");
    const size_t row_bytes = (size_t)st->W * st->C;
    if (pitch < row_bytes + (size_t)crop_x * st->C) return fail(CHB_ERR_INVALID, "chb_stack_upload: row pitch %zu too small", pitch);
    const uint8_t* src0 = host + (size_t)crop_y * pitch + (size_t)crop_x * st->C;  // pixel (0, 0) of the cropped frame
    int hs = -1;
    if (!pinned) {
        int rc = acquire_host_slot(st, hs);
        if (rc) return rc;
        uint8_t* dst = st->h_slot[hs];
        if (pitch == row_bytes) memcpy(dst, src0, row_bytes * (size_t)st->H);
        else
            for (int r = 0; r < st->H; r++) memcpy(dst + (size_t)r * row_bytes, src0 + (size_t)r * pitch, row_bytes);
        src0 = dst;
        pitch = row_bytes;
    }
    const int rc = stage_frame(st, frame_idx, [&](size_t bi, Band& b, Dev& d, uint8_t* dst) -> cudaError_t {
        const uint8_t* src = src0 + (size_t)b.row0 * pitch;
        cudaError_t e = pitch == row_bytes ? cudaMemcpyAsync(dst, src, b.frame_bytes, cudaMemcpyHostToDevice, d.copy)
                                           : cudaMemcpy2DAsync(dst, row_bytes, src, pitch, row_bytes, (size_t)b.rows, cudaMemcpyHostToDevice, d.copy);
        if (e == cudaSuccess && hs >= 0) e = cudaEventRecord(st->h_done[hs][bi], d.copy);
        return e;
    });
    if (hs >= 0) {
        std::lock_guard<std::mutex> lk(st->upload_mu);
        st->h_pending[hs] = true;
    }
    if (hs >= 0) release_host_slot(st, hs);
    return rc;
}

extern "C" int chb_stack_upload(chb_stack* st, int frame_idx, const uint8_t* host_pixels, size_t row_pitch, int crop_x, int crop_y) {
    return upload_impl(st, frame_idx, host_pixels, row_pitch, crop_x, crop_y, false);
}
extern "C" int chb_stack_upload_pinned(chb_stack* st, int frame_idx, const uint8_t* pinned_pixels, size_t row_pitch, int crop_x, int crop_y) {
    return upload_impl(st, frame_idx, pinned_pixels, row_pitch, crop_x, crop_y, true);
}

#include "chb_jpeg.inc"

extern "C" int chb_stack_download(chb_stack* st, int frame_idx, uint8_t* host_pixels, size_t row_pitch) {
    if (!st || !host_pixels) return fail(CHB_ERR_INVALID, "chb_stack_download: null argument");
    if (frame_idx < 0 || frame_idx >= st->N) return fail(CHB_ERR_INVALID, "chb_stack_download: frame %d outside [0, %d)", frame_idx, st->N);
    const size_t row_bytes = (size_t)st->W * st->C;
    if (row_pitch < row_bytes) return fail(CHB_ERR_INVALID, "chb_stack_download: row pitch %zu too small", row_pitch);
    std::lock_guard<std::mutex> lk(st->upload_mu);
    int rc = flush_ingest_locked(st);
    if (rc) return rc;
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(d.copy));
        CU(cudaStreamSynchronize(d.pack));
        if (!b.d_tmp) CU(cudaMalloc(&b.d_tmp, b.frame_bytes));
        unpack_frame_kernel<<<grid_for(b.n_pixels, 256, d.sm_count, 8), 256, 0, d.pack>>>(b.d_stack, b.d_tmp, b.n_pixels, st->C, st->NG, frame_idx);
        g_launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpy2DAsync(host_pixels + (size_t)b.row0 * row_pitch, row_pitch, b.d_tmp, row_bytes, row_bytes, (size_t)b.rows,
                             cudaMemcpyDeviceToHost, d.pack));
        CU(cudaStreamSynchronize(d.pack));
    }
    return CHB_OK;
}

extern "C" int chb_stack_sync(chb_stack* st) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_stack_sync: null stack");
    std::lock_guard<std::mutex> lk(st->upload_mu);
    int rc = flush_ingest_locked(st);
    if (rc) return rc;
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(d.copy));
        CU(cudaStreamSynchronize(d.pack));
    }
    return CHB_OK;
}

extern "C" int chb_stack_fill_synthetic(chb_stack* st, int kind, uint64_t seed, int row0_global, int full_height) {
    return chb_stack_fill_synthetic_blocks(st, kind, seed, row0_global, full_height, 0, 0);
}

// Same series for an interleaved row-block shard (chb_outlier_params.block_pixels): local row r of the stack is row
// row0_global + r + (r / block_rows) * block_skip_rows of the whole image.
extern "C" int chb_stack_fill_synthetic_blocks(chb_stack* st, int kind, uint64_t seed, int row0_global, int full_height, int block_rows,
                                               int block_skip_rows) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_stack_fill_synthetic: null stack");
    if (kind < 1 || kind > 4) return fail(CHB_ERR_INVALID, "chb_stack_fill_synthetic: unknown kind %d", kind);
    if (block_rows < 0 || block_skip_rows < 0 || (block_rows > 0 && st->bands.size() > 1))
        return fail(CHB_ERR_INVALID, "chb_stack_fill_synthetic_blocks: interleaved row blocks need a single-device stack");
    std::lock_guard<std::mutex> lk(st->upload_mu);
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        const long long units = b.n_tiles * st->C * st->NG * kTilePixels;
        synth_fill_kernel<<<grid_for(units, 256, d.sm_count, 16), 256, 0, d.pack>>>(b.d_stack, b.n_pixels, b.n_tiles, st->C, st->NG, st->N, kind,
                                                                                   (unsigned long long)seed, st->W, row0_global + b.row0, full_height,
                                                                                   block_rows, block_skip_rows);
        g_launches++;
        CU(cudaGetLastError());
    }
    for (Band& b : st->bands) {
        CU(cudaSetDevice(st->ctx->devs[b.dev_slot].id));
        CU(cudaStreamSynchronize(st->ctx->devs[b.dev_slot].pack));
    }
    std::fill(st->uploaded.begin(), st->uploaded.end(), 1);
    return CHB_OK;
}

extern "C" int chb_synth_frame_host(int kind, uint64_t seed, int frame_idx, int n_frames, int width, int full_height, int channels,
                                    int row0, int rows, uint8_t* out) {
    if (!out || kind < 1 || kind > 4 || width < 1 || rows < 0 || channels < 1 || channels > 4) return fail(CHB_ERR_INVALID, "chb_synth_frame_host: bad argument");
    for (int r = 0; r < rows; r++)
        for (int x = 0; x < width; x++)
            for (int c = 0; c < channels; c++)
                out[((size_t)r * width + x) * channels + c] = synth_byte(kind, seed, frame_idx, n_frames, row0 + r, x, c, width, full_height);
    return CHB_OK;
}

// ------------------------------------------------------------------------------------------------ window tables
struct Window {
    std::vector<int32_t> frames;  // position -> frame
    int g0 = 0, n_groups = 0;
};

static int build_window(chb_stack* st, const int32_t* indices, int n_indices, Window& w, const char* who) {
    if (indices) {
        if (n_indices < 1) return fail(CHB_ERR_INVALID, "%s: empty window", who);
        w.frames.assign(indices, indices + n_indices);
        for (int i = 0; i < n_indices; i++) {
            if (indices[i] < 0 || indices[i] >= st->N) return fail(CHB_ERR_INVALID, "%s: frame index %d outside [0, %d)", who, indices[i], st->N);
            if (i > 0 && indices[i] <= indices[i - 1]) return fail(CHB_ERR_INVALID, "%s: window indices must be strictly ascending (src/chrono.rs:113-139)", who);
        }
    } else {
        w.frames.resize(st->N);
        for (int i = 0; i < st->N; i++) w.frames[i] = i;
    }
    for (int f : w.frames)
        if (!st->uploaded[f]) return fail(CHB_ERR_STATE, "%s: frame %d of the window has not been uploaded", who, f);
    w.g0 = w.frames.front() / kGroupFrames;
    w.n_groups = w.frames.back() / kGroupFrames - w.g0 + 1;
    return CHB_OK;
}

static void byte_masks(const std::vector<int32_t>& frames, int g0, int cap_groups, uint32_t* out) {
    memset(out, 0, sizeof(uint32_t) * (size_t)cap_groups * 4);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(out);
    for (int f : frames) bytes[f - g0 * kGroupFrames] = 0xFF;
}

// deterministic replacement of rand::seq::sample_indices (src/chrono.rs:157): cnt distinct positions of [0, n), ascending
static void sample_positions(uint64_t seed, int n, int cnt, std::vector<int32_t>& out) {
    std::vector<int32_t> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int i = 0; i < cnt; i++) {
        int jx = i + (int)rng_range(seed ^ 0x5AFEC0DE5EEDULL, (uint64_t)i, 1, (uint32_t)(n - i));
        std::swap(perm[i], perm[jx]);
    }
    out.assign(perm.begin(), perm.begin() + cnt);
    std::sort(out.begin(), out.end());
}

extern "C" int chb_sample_positions(uint64_t seed, int n, int cnt, int32_t* out) {
    if (!out || n < 1 || cnt < 1 || cnt > n) return fail(CHB_ERR_INVALID, "chb_sample_positions: bad argument");
    std::vector<int32_t> v;
    sample_positions(seed, n, cnt, v);
    memcpy(out, v.data(), sizeof(int32_t) * (size_t)cnt);
    return CHB_OK;
}

// ------------------------------------------------------------------------------------------------ K1 dispatch

// Absolute thresholds with every weight 0 or 1: 4 * dist_sq is an exact integer below 2^20 and the outlier test becomes
// 4 * dist_sq >= ceil(4 * thr_sq) (see IntDist in chb_kernels.cuh).
static void set_int_dist(OutlierArgs& a) {
    bool ok = a.absolute && a.thr_sq == a.thr_sq && a.thr_sq >= 0.0f && a.thr_sq < 1.0e6f;
    for (int i = 0; i < a.C; i++) ok = ok && (a.w[i] == 0.0f || a.w[i] == 1.0f);
    a.int_dist = ok ? 1 : 0;
    a.thr4 = ok ? (int)ceilf(4.0f * a.thr_sq) : 0;
}
typedef void (*OutlierKernel)(const OutlierArgs);
struct Variant { int wpl, g; };
// capacity (frames) = 16 * wpl * g
static const Variant kVariants[] = {{1, 1}, {2, 1}, {4, 1}, {8, 1}, {13, 1}, {7, 2}, {8, 2}, {8, 4}, {16, 4}, {8, 8}, {8, 16}, {8, 32}};
static constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <int C, int GENERIC>
static OutlierKernel hard_kernel_for(int v) {
    switch (v) {
        case 0: return outlier_hard_kernel<C, 1, 1, GENERIC>;
        case 1: return outlier_hard_kernel<C, 2, 1, GENERIC>;
        case 2: return outlier_hard_kernel<C, 4, 1, GENERIC>;
        case 3: return outlier_hard_kernel<C, 8, 1, GENERIC>;
        case 4: return outlier_hard_kernel<C, 13, 1, GENERIC>;
        case 5: return outlier_hard_kernel<C, 7, 2, GENERIC>;
        case 6: return outlier_hard_kernel<C, 8, 2, GENERIC>;
        case 7: return outlier_hard_kernel<C, 8, 4, GENERIC>;
        case 8: return outlier_hard_kernel<C, 16, 4, GENERIC>;
        case 9: return outlier_hard_kernel<C, 8, 8, GENERIC>;
        case 10: return outlier_hard_kernel<C, 8, 16, GENERIC>;
        default: return outlier_hard_kernel<C, 8, 32, GENERIC>;
    }
}

template <int C, int GENERIC>  // GENERIC here is the kernel MODE: 0 generic, 1 absolute, 2 relative thresholds
static OutlierKernel kernel_for(int v) {
    switch (v) {
        case 0: return outlier_kernel<C, 1, 1, GENERIC>;
        case 1: return outlier_kernel<C, 2, 1, GENERIC>;
        case 2: return outlier_kernel<C, 4, 1, GENERIC>;
        case 3: return outlier_kernel<C, 8, 1, GENERIC>;
        case 4: return outlier_kernel<C, 13, 1, GENERIC>;
        case 5: return outlier_kernel<C, 7, 2, GENERIC>;
        case 6: return outlier_kernel<C, 8, 2, GENERIC>;
        case 7: return outlier_kernel<C, 8, 4, GENERIC>;
        case 8: return outlier_kernel<C, 16, 4, GENERIC>;
        case 9: return outlier_kernel<C, 8, 8, GENERIC>;
        case 10: return outlier_kernel<C, 8, 16, GENERIC>;
        default: return outlier_kernel<C, 8, 32, GENERIC>;
    }
}

static void quantile_ranks(int len, float q, int& r0, int& r1, float& frac) {  // src/chrono.rs:568-579
    float pos = (float)(len + 1) * q;
    int p1 = (int)pos - 1;
    float fr = pos - truncf(pos);
    if (fr < 0.001f) { r0 = r1 = p1; frac = 0.0f; }
    else if (fr > 0.999f) { r0 = r1 = p1 + 1; frac = 0.0f; }
    else { r0 = p1; r1 = p1 + 1; frac = fr; }
}

static int fade_to_dev(const chb_fade& f, FadeDev& out, const char* who) {
    out.is_none = f.is_none ? 1 : 0;
    out.mode = f.mode;
    out.absolute = f.absolute ? 1 : 0;
    out.offset = f.offset;
    out.n_values = f.n_values;
    out.values = nullptr;
    if (!f.is_none) {
        if (f.n_values < 1 || !f.values) return fail(CHB_ERR_INVALID, "%s: a fade needs at least one value", who);
        if (f.n_values > CHB_MAX_FADE_VALUES)  // a build limit, not invalid input: the reference accepts any length
            return fail(CHB_ERR_UNSUPPORTED, "%s: fade of %d values; this build holds at most %d", who, f.n_values, CHB_MAX_FADE_VALUES);
        if (f.mode != CHB_FADE_CLAMP && f.mode != CHB_FADE_REPEAT) return fail(CHB_ERR_INVALID, "%s: unknown fade mode", who);
    }
    return CHB_OK;
}

static int collect_outlier(chb_stack* st, int slot, const chb_debug_planes* dbg, float* kernel_ms, bool want_mask);

static int outlier_impl(chb_stack* st, int slot, const chb_outlier_params* prm, const int32_t* indices, int n_indices, bool want_mask,
                        const chb_debug_planes* dbg, float* kernel_ms, bool enqueue_only = false) {
    if (!st || !prm) return fail(CHB_ERR_INVALID, "chb_outlier: null argument");
    if (prm->background > CHB_BG_MEDIAN || prm->outlier > CHB_OUT_BACKWARD) return fail(CHB_ERR_INVALID, "chb_outlier: unknown background / outlier mode");
    if ((prm->block_pixels != 0 || prm->block_skip != 0) && (st->bands.size() > 1 || prm->block_pixels == 0))
        return fail(CHB_ERR_INVALID, "chb_outlier: interleaved row blocks need block_pixels > 0 and a single-device stack (one process per GPU)");
    // the caller owns call slot `slot`: launch and fetch form one critical section per slot
    CallSlot& cs = st->slots[slot];
    Window win;
    int rc = build_window(st, indices, n_indices, win, "chb_outlier");
    if (rc) return rc;
    rc = flush_ingest(st);  // frames uploaded but not yet re-laid-out (incomplete groups)
    if (rc) return rc;
    const int n = (int)win.frames.size();
    // --sample (src/chrono.rs:151-163)
    std::vector<int32_t> spos;
    bool sub = false;
    if (prm->sample_count >= 0) {
        if (prm->sample_count == 0) return fail(CHB_ERR_INVALID, "chb_outlier: --sample 0 (the reference panics on an empty sample)");
        if (prm->sample_count < n) {
            sample_positions(prm->seed, n, prm->sample_count, spos);
            sub = true;
        }
    }
    const int n_sub = sub ? (int)spos.size() : n;
    if (!prm->thr_absolute && n_sub < 3)
        return fail(CHB_ERR_INVALID, "chb_outlier: relative thresholds need at least 3 samples (quantile() underflows, src/chrono.rs:569-570)");
    int vidx = -1;
    for (int v = 0; v < kNumVariants; v++)
        if (kVariants[v].wpl * kVariants[v].g >= win.n_groups) { vidx = v; break; }
    {   // tuning aid: pick a larger-capacity variant by table index
        const int v = g_tune.force_variant.load();
        if (v >= 0 && v < kNumVariants && kVariants[v].wpl * kVariants[v].g >= win.n_groups) vidx = v;
    }
    // Series longer than the register-resident variants hold (more than 4096 frames): every pixel goes through the histogram tier
    // (outlier_hist_kernel reads a series of any length: one warp per pixel, a 256-bin histogram per band) and the per-frame path
    // for what its certificate cannot clear. Whole-stack launches only (a window or a --sample subset of such a series would need
    // byte masks the histogram kernel does not apply).
    const bool long_series = vidx < 0;
    if (long_series && (sub || n != st->N))
        return fail(CHB_ERR_UNSUPPORTED, "chb_outlier: a window or --sample subset spanning %d frames; beyond %d frames only whole-stack launches are supported",
                    win.n_groups * 16, kMaxWindowFrames);
    const Variant var = long_series ? Variant{0, 1} : kVariants[vidx];
    const int cap_groups = var.wpl * var.g;

    OutlierArgs a;
    memset(&a, 0, sizeof a);
    a.NG = st->NG; a.C = st->C;
    a.g0 = win.g0; a.n_groups = win.n_groups;
    a.n = n; a.n_sub = n_sub;
    a.first_frame = win.frames.front() & 15;
    a.inv_n_sub = 1.0f / (float)n_sub;
    // median ranks (src/chrono.rs:582-591)
    if ((n_sub + 1) % 2 == 0) a.rk[2] = a.rk[3] = (n_sub + 1) / 2 - 1;
    else { a.rk[2] = (n_sub + 1) / 2 - 1; a.rk[3] = (n_sub + 1) / 2; }
    if (!prm->thr_absolute) {
        quantile_ranks(n_sub, 0.25f, a.rk[0], a.rk[1], a.q1_frac);
        quantile_ranks(n_sub, 0.75f, a.rk[4], a.rk[5], a.q3_frac);
    }
    a.absolute = prm->thr_absolute ? 1 : 0;
    a.thr_min = prm->thr_min; a.thr_max = prm->thr_max; a.thr_scale = prm->thr_scale;
    a.thr_sq = prm->thr_min * prm->thr_min;  // src/chrono.rs:220
    for (int i = 0; i < 4; i++) a.w[i] = prm->weights[i];
    set_int_dist(a);
    a.bg = prm->background; a.om = prm->outlier;
    rc = fade_to_dev(prm->fade, a.fade, "chb_outlier");
    if (rc) return rc;
    a.frame_offset = indices ? indices[0] : 0;  // src/chrono.rs:102-103
    a.contig_f0 = (win.frames.back() - win.frames.front() + 1 == n) ? win.frames.front() : -1;
    a.exact_quartiles = (dbg && (dbg->q1 || dbg->q3)) ? 1 : 0;
    a.seed = prm->seed;
    // dense per-frame pass with per-group outlier masks: integer distance, contiguous window of at most kMaskGroups groups
    a.mask_path = (a.int_dist && a.thr4 >= 5 && a.contig_f0 >= 0 && win.n_groups <= kMaskGroups) ? 1 : 0;  // thr4 >= 5: see dense_masks_int
    // ... run inside the streaming kernel (lane = pixel: the G == 1 variants) for tiles with at least this many uncertified pixels
    a.inline_min = (a.mask_path && var.g == 1) ? g_tune.inline_min.load() : 0;
    a.hard_inline_min = (a.mask_path && var.g == 1) ? g_tune.hard_inline_min.load() : 0;
    a.hard_window = (var.g == 1 && !sub && g_tune.hard_window.load() != 0) ? 1 : 0;

    // host tables
    if (!long_series) byte_masks(win.frames, win.g0, cap_groups, cs.h_wmask);
    unsigned patch = 0;
    for (int i = 0; i < var.wpl; i++)
        for (int jx = 0; jx < var.g; jx++) {
            const uint32_t* m = cs.h_wmask + (size_t)(i * var.g + jx) * 4;
            if ((m[0] & m[1] & m[2] & m[3]) != 0xffffffffu && i < var.wpl - 1) patch |= 1u << i;
        }
    a.patch_slots = patch;
    a.lead_slots_full = (!long_series && win.n_groups >= (var.wpl - 1) * var.g) ? 1 : 0;
    // frames that exist in the stack, lie inside the span, but are not part of the window must be masked after the load
    {
        int in_span_existing = std::min(st->N, (win.g0 + win.n_groups) * kGroupFrames) - win.g0 * kGroupFrames;
        a.window_masked = (in_span_existing != n) ? 1 : 0;
    }
    if (sub) {
        std::vector<int32_t> sframes(spos.size());
        for (size_t i = 0; i < spos.size(); i++) sframes[i] = win.frames[spos[i]];
        byte_masks(sframes, win.g0, cap_groups, cs.h_smask);
    }
    memcpy(cs.h_win, win.frames.data(), sizeof(int32_t) * (size_t)n);
    if (!prm->fade.is_none) memcpy(cs.h_fade, prm->fade.values, sizeof(float) * (size_t)prm->fade.n_values);

    // the lean kernel covers whole-stack launches whose leading register slots are all real frame groups
    const bool generic = sub || a.window_masked || a.patch_slots || !a.lead_slots_full || a.exact_quartiles;
    const int kmode = generic ? 0 : (a.absolute ? 1 : 2);
    OutlierKernel kern = nullptr, hard_kern = nullptr;
    if (long_series) { /* no streaming kernel */ }
    else if (st->C == 3) kern = kmode == 0 ? kernel_for<3, 0>(vidx) : (kmode == 1 ? kernel_for<3, 1>(vidx) : kernel_for<3, 2>(vidx));
    else kern = kmode == 0 ? kernel_for<4, 0>(vidx) : (kmode == 1 ? kernel_for<4, 1>(vidx) : kernel_for<4, 2>(vidx));
    if (long_series) { /* the histogram kernel is the iterative tier */ }
    else if (st->C == 3) hard_kern = kmode == 0 ? hard_kernel_for<3, 0>(vidx) : (kmode == 1 ? hard_kernel_for<3, 1>(vidx) : hard_kernel_for<3, 2>(vidx));
    else hard_kern = kmode == 0 ? hard_kernel_for<4, 0>(vidx) : (kmode == 1 ? hard_kernel_for<4, 1>(vidx) : hard_kernel_for<4, 2>(vidx));

    // fingerprint of the device-side tables of this call
    std::vector<uint8_t> blob;
    {
        auto put = [&](const void* p, size_t nbytes) { const uint8_t* q = (const uint8_t*)p; blob.insert(blob.end(), q, q + nbytes); };
        const int hdr[4] = {1 /* outlier */, cap_groups, n, sub ? 1 : 0};
        put(hdr, sizeof hdr);
        put(cs.h_wmask, sizeof(uint32_t) * (size_t)cap_groups * 4);
        if (sub) put(cs.h_smask, sizeof(uint32_t) * (size_t)cap_groups * 4);
        put(cs.h_win, sizeof(int32_t) * (size_t)n);
        if (!prm->fade.is_none) put(cs.h_fade, sizeof(float) * (size_t)prm->fade.n_values);
    }
    const bool tables_cached = (blob == cs.last_tables);
    if (!tables_cached) cs.last_tables.clear();  // set again once every band's copies and launches were enqueued
    const size_t P = (size_t)st->W * st->H;
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        CallBuf& cb = b.call[slot];
        cudaStream_t s = call_stream(d, cb, slot);
        CU(wait_ingest(b, s));
        if (!tables_cached) {  // window / sample / fade tables: uploaded only when they differ from the previous call's
            CU(cudaMemcpyAsync(cb.d_wmask, cs.h_wmask, sizeof(uint32_t) * (size_t)cap_groups * 4, cudaMemcpyHostToDevice, s));
            if (sub) CU(cudaMemcpyAsync(cb.d_smask, cs.h_smask, sizeof(uint32_t) * (size_t)cap_groups * 4, cudaMemcpyHostToDevice, s));
            CU(cudaMemcpyAsync(cb.d_win, cs.h_win, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
            if (!prm->fade.is_none) CU(cudaMemcpyAsync(cb.d_fade, cs.h_fade, sizeof(float) * (size_t)prm->fade.n_values, cudaMemcpyHostToDevice, s));
        }
        if (!cb.d_queue) {  // one slot per pixel: the queues of the iterative tier and of the exact path cannot overflow
            CU(cudaMalloc(&cb.d_queue, sizeof(QueueEntry) * (size_t)b.n_pixels));
            CU(cudaMalloc(&cb.d_hqueue, sizeof(long long) * (size_t)b.n_pixels));
            // one allocation: words 0..7 the two queue counters, words 8..15 the call's four 64-bit counters, then the per-tile
            // flag words. The flag words are zeroed once here -- compact_hard_kernel clears every word it consumes -- so a call
            // starts with ONE 64-byte memset.
            CU(cudaMalloc(&cb.d_qcount, sizeof(uint32_t) * (32 + (size_t)b.n_tiles)));
            CU(cudaMemsetAsync(cb.d_qcount, 0, sizeof(uint32_t) * (32 + (size_t)b.n_tiles), s));
            cb.d_hflags = cb.d_qcount + 32;
            cb.qset = 0;
        }
        // The call's counters: two sets used alternately. The set of this call is zero already -- compact_hard_kernel of the
        // previous call cleared it -- so a call starts without a memset (one stream operation less between back-to-back calls);
        // only the launches without a compaction kernel (series beyond the register-resident variants) clear both sets here.
        if (long_series) CU(cudaMemsetAsync(cb.d_qcount, 0, sizeof(uint32_t) * 32, s));
        unsigned int* const qc = cb.d_qcount + 16 * cb.qset;
        unsigned int* const qc_other = cb.d_qcount + 16 * (cb.qset ^ 1);
        cb.last_qset = cb.qset;
        cb.qset ^= 1;
        OutlierArgs ab = a;
        ab.gq = cb.d_queue; ab.gq_count = qc;
        ab.ghq = cb.d_hqueue; ab.ghq_count = qc + 1;
        ab.hflags = cb.d_hflags;
        ab.stack = b.d_stack;
        ab.n_pixels = b.n_pixels; ab.n_tiles = b.n_tiles;
        ab.wmask = cb.d_wmask; ab.smask = sub ? cb.d_smask : nullptr; ab.win_frames = cb.d_win;
        ab.fade.values = cb.d_fade;
        ab.pixel_offset = prm->pixel_offset + (unsigned long long)b.row0 * st->W;
        ab.block_pixels = prm->block_pixels; ab.block_skip = prm->block_skip;
        ab.out_image = cb.d_out;
        ab.out_mask = want_mask ? cb.d_mask : nullptr;
        ab.counters = reinterpret_cast<unsigned long long*>(qc + 8);
        if (dbg) {
            if (dbg->median && !cb.d_dbg_median) CU(cudaMalloc(&cb.d_dbg_median, sizeof(float) * 4 * (size_t)b.n_pixels));
            if (dbg->q1 && !cb.d_dbg_q1) CU(cudaMalloc(&cb.d_dbg_q1, sizeof(float) * 4 * (size_t)b.n_pixels));
            if (dbg->q3 && !cb.d_dbg_q3) CU(cudaMalloc(&cb.d_dbg_q3, sizeof(float) * 4 * (size_t)b.n_pixels));
            if (dbg->n_outliers && !cb.d_dbg_nout) CU(cudaMalloc(&cb.d_dbg_nout, sizeof(int) * (size_t)b.n_pixels));
            // the median plane drives the other debug stores inside the kernel
            if (!cb.d_dbg_median && (dbg->q1 || dbg->q3)) CU(cudaMalloc(&cb.d_dbg_median, sizeof(float) * 4 * (size_t)b.n_pixels));
            if (cb.d_dbg_median) CU(cudaMemsetAsync(cb.d_dbg_median, 0, sizeof(float) * 4 * (size_t)b.n_pixels, s));
            if (cb.d_dbg_q1) CU(cudaMemsetAsync(cb.d_dbg_q1, 0, sizeof(float) * 4 * (size_t)b.n_pixels, s));
            if (cb.d_dbg_q3) CU(cudaMemsetAsync(cb.d_dbg_q3, 0, sizeof(float) * 4 * (size_t)b.n_pixels, s));
            ab.dbg_median = cb.d_dbg_median;
            ab.dbg_q1 = dbg->q1 ? cb.d_dbg_q1 : nullptr;
            ab.dbg_q3 = dbg->q3 ? cb.d_dbg_q3 : nullptr;
            ab.dbg_nout = dbg->n_outliers ? cb.d_dbg_nout : nullptr;
        }
        const long long n_tasks = b.n_tiles * var.g;
        if (n_tasks >= (1LL << 31)) return fail(CHB_ERR_UNSUPPORTED, "chb_outlier: band too large (%lld tile slices)", n_tasks);
        if (long_series) {  // histogram tier over every pixel of the band, then the per-frame path
            ab.hist_all = 1;
            CU(cudaEventRecord(cb.ev0, s));
            CU(cudaEventRecord(cb.ev_mid, s));
            if (st->C == 3) outlier_hist_kernel<3><<<d.sm_count * 8, kWarpsPerCta * 32, 0, s>>>(ab);
            else outlier_hist_kernel<4><<<d.sm_count * 8, kWarpsPerCta * 32, 0, s>>>(ab);
            if (st->C == 3) outlier_exact_kernel<3><<<d.sm_count * 4, 256, 0, s>>>(ab);
            else outlier_exact_kernel<4><<<d.sm_count * 4, 256, 0, s>>>(ab);
            g_launches += 2;
            CU(cudaGetLastError());
            CU(cudaEventRecord(cb.ev1, s));
            CU(cudaMemcpyAsync(cs.h_counters + 4 * b.dev_slot, qc + 8, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, s));
            continue;
        }
        int occ = 1;
        const int smem = outlier_smem_bytes(var.wpl, var.g, st->C);
        {
            int krc = kernel_config((const void*)kern, d.id, kWarpsPerCta * 32, smem, &occ);
            if (krc) return krc;
        }
        // persistent grid: every resident warp strides over the tile slices, so its exact-path queue fills up
        const int blocks = grid_for(n_tasks * 32, kWarpsPerCta * 32, d.sm_count, std::max(1, occ));
        CU(cudaEventRecord(cb.ev0, s));
        kern<<<blocks, kWarpsPerCta * 32, smem, s>>>(ab);
        CU(cudaEventRecord(cb.ev_mid, s));
        // iterative tier: long whole-stack series with relative thresholds (six ranks per band) use the histogram kernel -- one
        // shared-memory atomic per sample instead of the solver's repeated passes (measured on 1000 x UHD: 1.0 ms against
        // 1.7 ms; with absolute thresholds, two ranks, the solver's 0.6 ms wins). CHB_HIST=0 / 1 forces the choice (tests).
        compact_hard_kernel<<<std::min<long long>(d.sm_count * 4, (b.n_tiles + 1023) / 1024), 256, 0, s>>>(cb.d_hflags, b.n_tiles, cb.d_hqueue, qc + 1, qc_other);
        bool use_hist = kmode == 2 && n >= 256;
        if (g_tune.hist.load() >= 0) use_hist = kmode != 0 && n >= 256 && g_tune.hist.load() != 0;
        // with the dense per-frame pass in the iterative tier's kernel (every uncertified pixel of its warp-fulls finished in place)
        // that kernel also takes the streaming kernel's own queue and the exact-path launch is dropped
        ab.hard_drains_all = (!use_hist && ab.mask_path && var.g == 1 && kmode != 2 && ab.hard_inline_min == 1 && g_tune.hard_drains.load() != 0) ? 1 : 0;
        // the two tier kernels are programmatic dependent launches: their CTAs become resident while the previous kernel's
        // last CTAs drain and wait (griddepcontrol.wait) for its results, which hides two launch latencies per call
        const bool use_pdl = g_tune.pdl.load() != 0;
        auto launch_dep = [&](OutlierKernel k, int grid, int block, size_t sh) -> cudaError_t {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = sh; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
            return cudaLaunchKernelEx(&cfg, k, ab);
        };
        if (use_hist) {
            CU(launch_dep(st->C == 3 ? outlier_hist_kernel<3> : outlier_hist_kernel<4>, d.sm_count * 8, kWarpsPerCta * 32, 0));
        } else {
            {
                int krc = kernel_config((const void*)hard_kern, d.id, kWarpsPerCta * 32, smem, nullptr);
                if (krc) return krc;
            }
            CU(launch_dep(hard_kern, blocks, kWarpsPerCta * 32, smem));
        }
        if (!ab.hard_drains_all) CU(launch_dep(st->C == 3 ? outlier_exact_kernel<3> : outlier_exact_kernel<4>, d.sm_count * 4, 256, 0));
        g_launches += ab.hard_drains_all ? 3 : 4;
        CU(cudaGetLastError());
        CU(cudaEventRecord(cb.ev1, s));
        // (a call that is only enqueued leaves its counters on the device: whoever waits for it fetches them, and a run of
        // back-to-back calls has one stream operation less between the last kernel of a call and the first of the next)
        if (!(enqueue_only && tables_cached))
            CU(cudaMemcpyAsync(cs.h_counters + 4 * b.dev_slot, qc + 8, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, s));
    }
    cs.counters_pending = enqueue_only && tables_cached;
    (void)P;
    if (!tables_cached) cs.last_tables = blob;
    cs.last_kind = 1;
    if (enqueue_only && tables_cached) {  // nothing on the host is reused before the launch has read it: return without waiting
        cs.last_has_mask = want_mask;
        return CHB_OK;
    }
    return collect_outlier(st, slot, dbg, kernel_ms, want_mask);
}

// Waits for the launches of every band and gathers timing, counters and debug planes.
static int collect_outlier(chb_stack* st, int slot, const chb_debug_planes* dbg, float* kernel_ms, bool want_mask) {
    CallSlot& cs = st->slots[slot];
    float ms_max = 0.0f, main_max = 0.0f;
    uint64_t warnings = 0, slow = 0, hard = 0;
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CallBuf& cb = b.call[slot];
        CU(cudaSetDevice(d.id));
        if (cs.counters_pending)
            CU(cudaMemcpyAsync(cs.h_counters + 4 * b.dev_slot, cb.d_qcount + 16 * cb.last_qset + 8, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost,
                               call_stream(d, cb, slot)));
        CU(cudaStreamSynchronize(call_stream(d, cb, slot)));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, cb.ev0, cb.ev1));
        ms_max = std::max(ms_max, ms);
        {
            float mm = 0.0f;
            if (cudaEventElapsedTime(&mm, cb.ev0, cb.ev_mid) == cudaSuccess) main_max = std::max(main_max, mm);
            else cudaGetLastError();
        }
        warnings += cs.h_counters[4 * b.dev_slot];
        slow += cs.h_counters[4 * b.dev_slot + 1];
        hard += cs.h_counters[4 * b.dev_slot + 2];
        if (dbg) {
            const size_t off = (size_t)b.row0 * st->W;
            if (dbg->median) CU(cudaMemcpy(dbg->median + off * 4, cb.d_dbg_median, sizeof(float) * 4 * (size_t)b.n_pixels, cudaMemcpyDeviceToHost));
            if (dbg->q1) CU(cudaMemcpy(dbg->q1 + off * 4, cb.d_dbg_q1, sizeof(float) * 4 * (size_t)b.n_pixels, cudaMemcpyDeviceToHost));
            if (dbg->q3) CU(cudaMemcpy(dbg->q3 + off * 4, cb.d_dbg_q3, sizeof(float) * 4 * (size_t)b.n_pixels, cudaMemcpyDeviceToHost));
            if (dbg->n_outliers) CU(cudaMemcpy(dbg->n_outliers + off, cb.d_dbg_nout, sizeof(int) * (size_t)b.n_pixels, cudaMemcpyDeviceToHost));
        }
    }
    if (kernel_ms) *kernel_ms = ms_max;
    cs.counters_pending = false;
    cs.last_has_mask = want_mask;
    cs.last_warnings = warnings;
    g_last_slow = slow;
    g_last_hard = hard;
    g_last_main_ms = main_max;
    return CHB_OK;
}

static int fetch_impl(chb_stack* st, int slot, uint8_t* out_image, uint8_t* out_mask, uint64_t* n_warnings) {
    CallSlot& cs = st->slots[slot];
    if (cs.last_kind != 1) return fail(CHB_ERR_STATE, "chb_fetch_last: %s", cs.last_kind == 2 ? "the last call was a video run (its planes went to the caller's buffers)" : "no call has run on this stack yet");
    if (out_mask && !cs.last_has_mask) return fail(CHB_ERR_STATE, "chb_fetch_last: the last call did not produce a mask");
    if (cs.counters_pending) {  // the last call was only enqueued: its counters (the warning count) are still on the devices
        int rc = collect_outlier(st, slot, nullptr, nullptr, cs.last_has_mask);
        if (rc) return rc;
    }
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CallBuf& cb = b.call[slot];
        cudaStream_t s = call_stream(d, cb, slot);
        CU(cudaSetDevice(d.id));
        const size_t off = (size_t)b.row0 * st->W * st->C;
        if (out_image) CU(cudaMemcpyAsync(out_image + off, cb.d_out, b.frame_bytes, cudaMemcpyDeviceToHost, s));
        if (out_mask) CU(cudaMemcpyAsync(out_mask + off, cb.d_mask, b.frame_bytes, cudaMemcpyDeviceToHost, s));
    }
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(call_stream(d, b.call[slot], slot)));
    }
    if (n_warnings) *n_warnings = cs.last_warnings;
    return CHB_OK;
}

extern "C" int chb_outlier_debug(chb_stack* st, const chb_outlier_params* prm, const int32_t* indices, int n_indices, uint8_t* out_image,
                                 uint8_t* out_mask, uint64_t* n_warnings, const chb_debug_planes* dbg) {
    if (!out_image || !st) return fail(CHB_ERR_INVALID, "chb_outlier: null argument");
    int rc = chb_stack_sync(st);
    if (rc) return rc;
    SlotGuard g(st, false);  // any free call slot: concurrent callers overlap launches, tier kernels and D2H copies
    rc = ensure_call_slot(st, g.slot);
    if (rc) return rc;
    rc = outlier_impl(st, g.slot, prm, indices, n_indices, out_mask != nullptr, dbg, nullptr);
    if (rc) return rc;
    return fetch_impl(st, g.slot, out_image, out_mask, n_warnings);
}
extern "C" int chb_outlier(chb_stack* st, const chb_outlier_params* prm, const int32_t* indices, int n_indices, uint8_t* out_image,
                           uint8_t* out_mask, uint64_t* n_warnings) {
    return chb_outlier_debug(st, prm, indices, n_indices, out_image, out_mask, n_warnings, nullptr);
}
// ---- device-side API: defined on "the last call" of the stack, i.e. on call slot 0
extern "C" int chb_outlier_device(chb_stack* st, const chb_outlier_params* prm, const int32_t* indices, int n_indices, int want_mask, float* kernel_ms) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_outlier_device: null stack");
    SlotGuard g(st, true);
    return outlier_impl(st, 0, prm, indices, n_indices, want_mask != 0, nullptr, kernel_ms);
}
extern "C" int chb_outlier_enqueue(chb_stack* st, const chb_outlier_params* prm, const int32_t* indices, int n_indices, int want_mask) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_outlier_enqueue: null stack");
    SlotGuard g(st, true);
    return outlier_impl(st, 0, prm, indices, n_indices, want_mask != 0, nullptr, nullptr, true);
}
extern "C" int chb_stack_wait(chb_stack* st, float* last_kernel_ms, uint64_t* n_warnings) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_stack_wait: null stack");
    SlotGuard g(st, true);
    if (st->slots[0].last_kind != 1) return fail(CHB_ERR_STATE, "chb_stack_wait: no single-window call is pending on this stack");
    int rc = collect_outlier(st, 0, nullptr, last_kernel_ms, st->slots[0].last_has_mask);
    if (rc) return rc;
    if (n_warnings) *n_warnings = st->slots[0].last_warnings;
    return CHB_OK;
}
extern "C" int chb_fetch_last_device(chb_stack* st, int dev_slot, void* d_image, void* d_mask) {
    if (!st || !d_image) return fail(CHB_ERR_INVALID, "chb_fetch_last_device: null argument");
    if (dev_slot < 0 || dev_slot >= (int)st->bands.size()) return fail(CHB_ERR_INVALID, "chb_fetch_last_device: bad device slot");
    SlotGuard g(st, true);
    if (st->slots[0].last_kind != 1) return fail(CHB_ERR_STATE, "chb_fetch_last_device: the last call on this stack was not a single-window call");
    if (d_mask && !st->slots[0].last_has_mask) return fail(CHB_ERR_STATE, "chb_fetch_last_device: the last call did not produce a mask");
    Band& b = st->bands[dev_slot];
    Dev& d = st->ctx->devs[b.dev_slot];
    CU(cudaSetDevice(d.id));
    CU(cudaMemcpyAsync(d_image, b.call[0].d_out, b.frame_bytes, cudaMemcpyDeviceToDevice, d.compute));
    if (d_mask) CU(cudaMemcpyAsync(d_mask, b.call[0].d_mask, b.frame_bytes, cudaMemcpyDeviceToDevice, d.compute));
    return CHB_OK;
}
extern "C" int chb_fetch_last(chb_stack* st, uint8_t* out_image, uint8_t* out_mask, uint64_t* n_warnings) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_fetch_last: null stack");
    SlotGuard g(st, true);
    return fetch_impl(st, 0, out_image, out_mask, n_warnings);
}

// ------------------------------------------------------------------------------------------------ K3 dispatch (chrono-video)
typedef void (*VideoKernel)(VideoArgs);
template <int C>
static VideoKernel video_kernel_for(int nw) {
    switch (nw) {
        case 2: return video_kernel<C, 2>;
        case 4: return video_kernel<C, 4>;
        case 6: return video_kernel<C, 6>;
        case 8: return video_kernel<C, 8>;
        case 10: return video_kernel<C, 10>;
        case 12: return video_kernel<C, 12>;
        case 14: return video_kernel<C, 14>;
        default: return video_kernel<C, 16>;
    }
}
static constexpr int kMaxVideoWindow = 64;
static constexpr unsigned int kVideoQueueEntries = 8u << 20;  // 512 MB; pixel-windows beyond it are finished inside video_kernel

// A run of n_windows windows of window_len consecutive frames, window i starting at frame first_start + i (the windows
// create_video builds for `--video-in a/b/1`, src/main.rs:262-286). Host outputs (may be null: results stay on the devices,
// only the last chunk's) are [n_windows][H*W*C] planes.
static int video_impl(chb_stack* st, const chb_outlier_params* prm, int first_start, int window_len, int n_windows, uint8_t* out_images,
                      uint8_t* out_masks, bool want_mask, uint64_t* warnings, float* kernel_ms) {
    if (!st || !prm) return fail(CHB_ERR_INVALID, "chb_outlier_video: null argument");
    if (prm->background > CHB_BG_MEDIAN || prm->outlier > CHB_OUT_BACKWARD) return fail(CHB_ERR_INVALID, "chb_outlier_video: unknown background / outlier mode");
    const int n = window_len;
    if (n < 1 || n_windows < 1 || first_start < 0 || (long long)first_start + n_windows - 1 + n > st->N)
        return fail(CHB_ERR_INVALID, "chb_outlier_video: windows [%d + i, %d + i + %d), i < %d leave the stack of %d frames", first_start, first_start, n,
                    n_windows, st->N);
    if (n > kMaxVideoWindow)
        return fail(CHB_ERR_UNSUPPORTED, "chb_outlier_video: windows of %d frames; the sliding kernel holds at most %d (use chb_outlier per window)", n, kMaxVideoWindow);
    if (prm->sample_count >= 0 && prm->sample_count < n)
        return fail(CHB_ERR_UNSUPPORTED, "chb_outlier_video: --sample below the window length is not supported by the sliding kernel (use chb_outlier per window)");
    if (!prm->thr_absolute && n < 3)
        return fail(CHB_ERR_INVALID, "chb_outlier_video: relative thresholds need at least 3 samples (quantile() underflows, src/chrono.rs:569-570)");
    for (int f = first_start; f < first_start + n_windows - 1 + n; f++)
        if (!st->uploaded[f]) return fail(CHB_ERR_STATE, "chb_outlier_video: frame %d was never uploaded", f);
    {
        int frc = flush_ingest(st);
        if (frc) return frc;
    }

    VideoArgs va;
    memset(&va, 0, sizeof va);
    OutlierArgs& a = va.o;
    a.NG = st->NG; a.C = st->C;
    a.n = n; a.n_sub = n;
    a.inv_n_sub = 1.0f / (float)n;
    if ((n + 1) % 2 == 0) a.rk[2] = a.rk[3] = (n + 1) / 2 - 1;  // median ranks (src/chrono.rs:582-591)
    else { a.rk[2] = (n + 1) / 2 - 1; a.rk[3] = (n + 1) / 2; }
    if (!prm->thr_absolute) {
        quantile_ranks(n, 0.25f, a.rk[0], a.rk[1], a.q1_frac);
        quantile_ranks(n, 0.75f, a.rk[4], a.rk[5], a.q3_frac);
    }
    a.absolute = prm->thr_absolute ? 1 : 0;
    a.thr_min = prm->thr_min; a.thr_max = prm->thr_max; a.thr_scale = prm->thr_scale;
    a.thr_sq = prm->thr_min * prm->thr_min;  // src/chrono.rs:220
    for (int i = 0; i < 4; i++) a.w[i] = prm->weights[i];
    set_int_dist(a);
    a.bg = prm->background; a.om = prm->outlier;
    int rc = fade_to_dev(prm->fade, a.fade, "chb_outlier_video");
    if (rc) return rc;
    a.seed = prm->seed;
    const int nw = std::max(2, 2 * (((n + 3) / 4 + 1) / 2));  // window words, rounded up to an even count
    auto word_mask = [&](int q) {
        uint32_t m = 0;
        for (int b = 0; b < 4; b++)
            if (4 * q + b < n) m |= 0xffu << (8 * b);
        return m;
    };
    va.mask_a = word_mask(nw - 2);
    va.mask_b = word_mask(nw - 1);
    va.res_words = (!prm->thr_absolute || prm->background == CHB_BG_AVERAGE) ? 2 : 1;
    const int smem = video_smem_bytes(st->C, va.res_words);
    VideoKernel kern = st->C == 3 ? video_kernel_for<3>(nw) : video_kernel_for<4>(nw);
    if (!prm->fade.is_none) memcpy(st->slots[0].h_fade, prm->fade.values, sizeof(float) * (size_t)prm->fade.n_values);
    st->slots[0].last_tables.clear();  // the fade table on the devices no longer belongs to a single-window call

    // chunks of whole 16-start blocks; two slots of output planes per band (kernel of chunk k+1 overlaps the D2H of chunk k)
    size_t max_frame_bytes = 0;
    for (Band& b : st->bands) max_frame_bytes = std::max(max_frame_bytes, b.frame_bytes);
    const int blocks_per_chunk = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)4 << 30) / (64 * max_frame_bytes)));
    const int chunk_windows = blocks_per_chunk * kVideoBlock;
    const int blk_first = first_start / kVideoBlock, blk_last = (first_start + n_windows - 1) / kVideoBlock;
    const int n_chunks = (blk_last - blk_first + blocks_per_chunk) / blocks_per_chunk;
    const size_t P = (size_t)st->W * st->H;

    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        if (b.v_cap_windows < chunk_windows) {
            for (int s = 0; s < 2; s++) {
                cudaFree(b.d_vout[s]); cudaFree(b.d_vmask[s]);
                b.d_vout[s] = b.d_vmask[s] = nullptr;
                CU(cudaMalloc(&b.d_vout[s], b.frame_bytes * (size_t)chunk_windows));
                CU(cudaMalloc(&b.d_vmask[s], b.frame_bytes * (size_t)chunk_windows));
                if (!b.v_done[s]) CU(cudaEventCreateWithFlags(&b.v_done[s], cudaEventDisableTiming));
                if (!b.v_copied[s]) CU(cudaEventCreateWithFlags(&b.v_copied[s], cudaEventDisableTiming));
            }
            b.v_cap_windows = chunk_windows;
        }
        if (!b.d_vqueue) {
            CU(cudaMalloc(&b.d_vqueue, sizeof(VideoQueueEntry) * (size_t)kVideoQueueEntries));
            CU(cudaMalloc(&b.d_vqcount, 2 * sizeof(unsigned int)));
        }
        if (b.vwarn_cap < n_windows) {
            cudaFree(b.d_vwarn);
            b.d_vwarn = nullptr;
            CU(cudaMalloc(&b.d_vwarn, sizeof(unsigned long long) * (size_t)n_windows));
            b.vwarn_cap = n_windows;
        }
        cudaStream_t s = d.compute;
        CU(wait_ingest(b, s));
        if (!prm->fade.is_none) CU(cudaMemcpyAsync(b.call[0].d_fade, st->slots[0].h_fade, sizeof(float) * (size_t)prm->fade.n_values, cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync(b.call[0].d_counters, 0, sizeof(unsigned long long) * 4, s));
        CU(cudaMemsetAsync(b.d_vwarn, 0, sizeof(unsigned long long) * (size_t)n_windows, s));
        CU(cudaEventRecord(b.call[0].ev0, s));
    }
    std::vector<bool> slot_in_flight(2 * st->bands.size(), false);
    for (int k = 0; k < n_chunks; k++) {
        const int slot = k & 1;
        const int cb0 = blk_first + k * blocks_per_chunk;
        const int cb1 = std::min(blk_last, cb0 + blocks_per_chunk - 1);
        const int s_lo = std::max(first_start, cb0 * kVideoBlock);                       // first window start of this chunk
        const int s_hi = std::min(first_start + n_windows - 1, cb1 * kVideoBlock + kVideoBlock - 1);
        const int w_lo = s_lo - first_start, cw = s_hi - s_lo + 1;
        for (size_t bi = 0; bi < st->bands.size(); bi++) {
            Band& b = st->bands[bi];
            Dev& d = st->ctx->devs[b.dev_slot];
            CU(cudaSetDevice(d.id));
            cudaStream_t s = d.compute;
            if (slot_in_flight[2 * bi + slot]) CU(cudaStreamWaitEvent(s, b.v_copied[slot], 0));  // the slot's previous planes have left
            VideoArgs vb = va;
            vb.o.stack = b.d_stack;
            vb.o.n_pixels = b.n_pixels; vb.o.n_tiles = b.n_tiles;
            vb.o.fade.values = b.call[0].d_fade;
            vb.o.pixel_offset = prm->pixel_offset + (unsigned long long)b.row0 * st->W;
            vb.o.block_pixels = prm->block_pixels; vb.o.block_skip = prm->block_skip;
            vb.o.counters = b.call[0].d_counters;
            vb.first_start = s_lo; vb.n_windows = cw;
            vb.blk0 = cb0; vb.n_blocks = cb1 - cb0 + 1;
            vb.out_stride = (long long)b.frame_bytes;
            vb.out_images = b.d_vout[slot];
            vb.out_masks = want_mask ? b.d_vmask[slot] : nullptr;
            vb.win_warnings = b.d_vwarn + w_lo;
            vb.gq = b.d_vqueue; vb.gq_count = b.d_vqcount; vb.gq_cap = kVideoQueueEntries;
            if (g_tune.video_queue_cap.load() >= 0)  // test aid: a small capacity exercises the in-place fallback
                vb.gq_cap = (unsigned int)std::min<long long>(kVideoQueueEntries, g_tune.video_queue_cap.load());
            CU(cudaMemsetAsync(b.d_vqcount, 0, sizeof(unsigned int), s));
            CU(cudaMemsetAsync(b.d_vqcount + 1, 0xff, sizeof(unsigned int), s));
            const long long n_tasks = b.n_tiles * vb.n_blocks;
            if (n_tasks >= (1LL << 31)) return fail(CHB_ERR_UNSUPPORTED, "chb_outlier_video: band too large (%lld tasks)", n_tasks);
            int occ = 1;
            {
                int krc = kernel_config((const void*)kern, d.id, kVideoWarps * 32, smem, &occ);
                if (krc) return krc;
            }
            const int blocks = grid_for(n_tasks * 32, kVideoWarps * 32, d.sm_count, std::max(1, occ));
            kern<<<blocks, kVideoWarps * 32, smem, s>>>(vb);
            if (st->C == 3) video_exact_kernel<3><<<d.sm_count * 8, 128, 0, s>>>(vb);
            else video_exact_kernel<4><<<d.sm_count * 8, 128, 0, s>>>(vb);
            g_launches += 2;
            CU(cudaGetLastError());
            if (out_images) {
                CU(cudaEventRecord(b.v_done[slot], s));
                CU(cudaStreamWaitEvent(d.copy, b.v_done[slot], 0));
                const size_t off = ((size_t)w_lo * P + (size_t)b.row0 * st->W) * st->C;
                CU(cudaMemcpy2DAsync(out_images + off, P * st->C, b.d_vout[slot], b.frame_bytes, b.frame_bytes, (size_t)cw, cudaMemcpyDeviceToHost, d.copy));
                if (out_masks)
                    CU(cudaMemcpy2DAsync(out_masks + off, P * st->C, b.d_vmask[slot], b.frame_bytes, b.frame_bytes, (size_t)cw, cudaMemcpyDeviceToHost, d.copy));
                CU(cudaEventRecord(b.v_copied[slot], d.copy));
                slot_in_flight[2 * bi + slot] = true;
            }
        }
    }
    float ms_max = 0.0f;
    uint64_t slow = 0, hard = 0, warn_total = 0;
    std::vector<unsigned long long> wtmp((size_t)n_windows);
    if (warnings) memset(warnings, 0, sizeof(uint64_t) * (size_t)n_windows);
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        CU(cudaEventRecord(b.call[0].ev1, d.compute));
        CU(cudaMemcpyAsync(st->slots[0].h_counters + 4 * b.dev_slot, b.call[0].d_counters, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, d.compute));
        CU(cudaStreamSynchronize(d.compute));
        CU(cudaStreamSynchronize(d.copy));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, b.call[0].ev0, b.call[0].ev1));
        ms_max = std::max(ms_max, ms);
        slow += st->slots[0].h_counters[4 * b.dev_slot + 1];
        hard += st->slots[0].h_counters[4 * b.dev_slot + 2];
        CU(cudaMemcpy(wtmp.data(), b.d_vwarn, sizeof(unsigned long long) * (size_t)n_windows, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n_windows; i++) {
            warn_total += wtmp[i];
            if (warnings) warnings[i] += wtmp[i];
        }
    }
    if (kernel_ms) *kernel_ms = ms_max;
    st->slots[0].last_warnings = warn_total;
    st->slots[0].last_kind = 2;
    g_last_slow = slow;
    g_last_hard = hard;
    return CHB_OK;
}

extern "C" int chb_outlier_video(chb_stack* st, const chb_outlier_params* prm, int first_start, int window_len, int n_windows, uint8_t* out_images,
                                 uint8_t* out_masks, uint64_t* n_warnings) {
    if (!st || !out_images) return fail(CHB_ERR_INVALID, "chb_outlier_video: null argument");
    int rc = chb_stack_sync(st);
    if (rc) return rc;
    SlotGuard g(st, true);
    return video_impl(st, prm, first_start, window_len, n_windows, out_images, out_masks, out_masks != nullptr, n_warnings, nullptr);
}
extern "C" int chb_outlier_video_device(chb_stack* st, const chb_outlier_params* prm, int first_start, int window_len, int n_windows, int want_mask,
                                        float* kernel_ms) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_outlier_video_device: null stack");
    SlotGuard g(st, true);
    return video_impl(st, prm, first_start, window_len, n_windows, nullptr, nullptr, want_mask != 0, nullptr, kernel_ms);
}

// ------------------------------------------------------------------------------------------------ K2 dispatch
static int simple_impl(chb_stack* st, int slot, const chb_simple_params* prm, const int32_t* indices, int n_indices, float* kernel_ms) {
    if (!st || !prm) return fail(CHB_ERR_INVALID, "chb_simple: null argument");
    // the caller owns call slot `slot`
    CallSlot& cs = st->slots[slot];
    Window win;
    // SimpleProcessor accepts any index order in principle (src/simple.rs:142-146), but every caller passes ascending
    // windows (src/main.rs:398-404); the time-sliced stack relies on it.
    int rc = build_window(st, indices, n_indices, win, "chb_simple");
    if (rc) return rc;
    rc = flush_ingest(st);
    if (rc) return rc;
    const int n = (int)win.frames.size();
    SimpleArgs a;
    memset(&a, 0, sizeof a);
    a.NG = st->NG; a.C = st->C;
    a.g0 = win.g0; a.n_groups = win.n_groups;
    a.n = n;
    a.darker = prm->darker ? 1 : 0;
    for (int i = 0; i < 4; i++) a.w[i] = prm->weights[i];
    rc = fade_to_dev(prm->fade, a.fade, "chb_simple");
    if (rc) return rc;
    a.frame_offset = indices ? indices[0] : 0;  // src/simple.rs:54-57
    const bool fade = !prm->fade.is_none;
    // integer kernel: no fade and every weight of an existing band is exactly 0 or 1 (the CLI default is all ones)
    bool int_path = !fade;
    for (int i = 0; i < st->C; i++) {
        if (prm->weights[i] == 1.0f) a.use_mask |= 1u << i;
        else if (prm->weights[i] != 0.0f) int_path = false;
    }
    cs.last_tables.clear();  // this call overwrites the shared table buffers
    std::vector<uint32_t> masks((size_t)win.n_groups * 4);
    byte_masks(win.frames, win.g0, win.n_groups, masks.data());
    bool all_in = true;
    for (uint32_t m : masks) all_in = all_in && (m == 0xffffffffu);
    std::vector<int32_t> posg(win.n_groups, 0);
    {
        size_t k = 0;
        for (int gi = 0; gi < win.n_groups; gi++) {
            while (k < win.frames.size() && win.frames[k] < (win.g0 + gi) * kGroupFrames) k++;
            posg[gi] = (int32_t)k;
        }
    }
    if (!prm->fade.is_none) memcpy(cs.h_fade, prm->fade.values, sizeof(float) * (size_t)prm->fade.n_values);
    memcpy(cs.h_posg, posg.data(), sizeof(int32_t) * posg.size());
    std::vector<uint32_t*> tmp_masks;
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CU(cudaSetDevice(d.id));
        CallBuf& cb = b.call[slot];
        cudaStream_t s = call_stream(d, cb, slot);
        CU(wait_ingest(b, s));
        SimpleArgs ab = a;
        ab.stack = b.d_stack;
        ab.n_pixels = b.n_pixels; ab.n_tiles = b.n_tiles;
        ab.wmask = nullptr;
        if (!all_in) {
            uint32_t* dm = nullptr;
            if (win.n_groups <= kMaxGroupsTable) dm = cb.d_wmask;
            else { CU(cudaMalloc(&dm, sizeof(uint32_t) * masks.size())); tmp_masks.push_back(dm); }
            CU(cudaMemcpyAsync(dm, masks.data(), sizeof(uint32_t) * masks.size(), cudaMemcpyHostToDevice, s));
            ab.wmask = dm;
        }
        CU(cudaMemcpyAsync(cb.d_posg, cs.h_posg, sizeof(int32_t) * posg.size(), cudaMemcpyHostToDevice, s));
        ab.pos_of_group = cb.d_posg;
        if (fade) CU(cudaMemcpyAsync(cb.d_fade, cs.h_fade, sizeof(float) * (size_t)prm->fade.n_values, cudaMemcpyHostToDevice, s));
        ab.fade.values = cb.d_fade;
        ab.out_image = cb.d_out;
        const int blocks = grid_for(b.n_tiles * kTilePixels, 256, d.sm_count, 64);
        CU(cudaEventRecord(cb.ev0, s));
        if (int_path) {
            if (st->C == 3) simple_int_kernel<3><<<blocks, 256, 0, s>>>(ab);
            else simple_int_kernel<4><<<blocks, 256, 0, s>>>(ab);
        } else if (st->C == 3) {
            if (fade) simple_kernel<3, true><<<blocks, 256, 0, s>>>(ab);
            else simple_kernel<3, false><<<blocks, 256, 0, s>>>(ab);
        } else {
            if (fade) simple_kernel<4, true><<<blocks, 256, 0, s>>>(ab);
            else simple_kernel<4, false><<<blocks, 256, 0, s>>>(ab);
        }
        g_launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(cb.ev1, s));
    }
    float ms_max = 0.0f;
    for (Band& b : st->bands) {
        Dev& d = st->ctx->devs[b.dev_slot];
        CallBuf& cb = b.call[slot];
        CU(cudaSetDevice(d.id));
        CU(cudaStreamSynchronize(call_stream(d, cb, slot)));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, cb.ev0, cb.ev1));
        ms_max = std::max(ms_max, ms);
    }
    for (uint32_t* p : tmp_masks) cudaFree(p);
    if (kernel_ms) *kernel_ms = ms_max;
    cs.last_kind = 1;
    cs.last_has_mask = false;
    cs.last_warnings = 0;
    return CHB_OK;
}

extern "C" int chb_simple(chb_stack* st, const chb_simple_params* prm, const int32_t* indices, int n_indices, uint8_t* out_image) {
    if (!out_image || !st) return fail(CHB_ERR_INVALID, "chb_simple: null argument");
    int rc = chb_stack_sync(st);
    if (rc) return rc;
    SlotGuard g(st, false);
    rc = ensure_call_slot(st, g.slot);
    if (rc) return rc;
    rc = simple_impl(st, g.slot, prm, indices, n_indices, nullptr);
    if (rc) return rc;
    return fetch_impl(st, g.slot, out_image, nullptr, nullptr);
}
extern "C" int chb_simple_device(chb_stack* st, const chb_simple_params* prm, const int32_t* indices, int n_indices, float* kernel_ms) {
    if (!st) return fail(CHB_ERR_INVALID, "chb_simple_device: null stack");
    SlotGuard g(st, true);
    return simple_impl(st, 0, prm, indices, n_indices, kernel_ms);
}

// ------------------------------------------------------------------------------------------------ host arithmetic
extern "C" void chb_threshold_new(int absolute, float min, float max, float* out_min, float* out_max, float* out_scale) {
    // Threshold::new, src/options.rs:197-213 (f32 throughout)
    if (absolute) {
        *out_min = min * 255.0f;
        *out_max = max * 255.0f;
        *out_scale = 1.0f / ((max - min) * 255.0f);
    } else {
        *out_min = min;
        *out_max = max;
        *out_scale = 1.0f / (max - min);
    }
}

extern "C" int chb_fade_build(const int32_t* frames, const float* values, int n_pairs, float* out_values, int out_cap, int32_t* out_offset) {
    // Fade::new, src/options.rs:69-94
    if (!frames || !values || !out_values || !out_offset) return -1;
    if (n_pairs < 2) return -1;  // "Fade requires at least two frames specified."
    const int32_t offset = frames[0];
    const int32_t len = frames[n_pairs - 1] - offset;
    if (len < 0 || len + 1 > out_cap) return -1;
    int idx = 0;
    for (int32_t i = 0; i <= len; i++) {
        if (idx + 1 >= n_pairs) return -1;
        const int32_t f1 = frames[idx], f2 = frames[idx + 1];
        const float v1 = values[idx], v2 = values[idx + 1];
        const int32_t frame = i + offset;
        out_values[i] = v1 + (v2 - v1) * (float)(frame - f1) / (float)(f2 - f1);
        if (frame == f2 && idx + 2 < n_pairs) idx++;
    }
    *out_offset = offset;
    return len + 1;
}

// ------------------------------------------------------------------------------------------------ K4 dispatch (shake analysis)
struct chb_shake {
    chb_ctx* ctx = nullptr;
    int W = 0, H = 0, C = 0, n_anchors = 0, radius = 0, search = 0, size = 0, psize = 0, search_size = 0;
    std::vector<int32_t> anchors;
    uint8_t *d_windows = nullptr, *d_patches = nullptr, *h_patches = nullptr;
    int32_t *d_diffs = nullptr, *d_result = nullptr, *h_diffs = nullptr, *h_result = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;
};

extern "C" int chb_shake_destroy(chb_shake* sh) {
    if (!sh) return CHB_OK;
    cudaSetDevice(sh->ctx->devs[0].id);
    if (sh->stream) { cudaStreamSynchronize(sh->stream); cudaStreamDestroy(sh->stream); }
    cudaFree(sh->d_windows); cudaFree(sh->d_patches); cudaFree(sh->d_diffs); cudaFree(sh->d_result);
    if (sh->h_patches) cudaFreeHost(sh->h_patches);
    if (sh->h_diffs) cudaFreeHost(sh->h_diffs);
    if (sh->h_result) cudaFreeHost(sh->h_result);
    delete sh;
    return CHB_OK;
}

// ShakeAnalyzer::analyze up to the first frame (src/shake.rs:222-246): fill_windows (:307-336) on the host side of the
// boundary (a gather of n_anchors * (2r+1)^2 pixels), uploaded once.
extern "C" int chb_shake_create(chb_ctx* ctx, int width, int height, int channels, const int32_t* anchors_xy, int n_anchors, int anchor_radius,
                                int search_radius, const uint8_t* first_frame, size_t row_pitch, chb_shake** out) {
    if (!ctx || !anchors_xy || !first_frame || !out) return fail(CHB_ERR_INVALID, "chb_shake_create: null argument");
    if (width < 1 || height < 1 || channels < 1 || channels > 4 || n_anchors < 1 || anchor_radius < 0 || search_radius < 0)
        return fail(CHB_ERR_INVALID, "chb_shake_create: bad geometry (%dx%dx%d, %d anchors, radii %d/%d)", width, height, channels, n_anchors, anchor_radius,
                    search_radius);
    if (row_pitch < (size_t)width * channels) return fail(CHB_ERR_INVALID, "chb_shake_create: row pitch smaller than a row");
    if ((long long)(2 * (anchor_radius + search_radius) + 1) > 8192) return fail(CHB_ERR_UNSUPPORTED, "chb_shake_create: radii too large");
    for (int i = 0; i < n_anchors; i++) {
        const int cx = anchors_xy[2 * i], cy = anchors_xy[2 * i + 1];
        for (int k = 0; k < 4; k++) {
            const int xx = cx + ((k & 1) ? anchor_radius : -anchor_radius), yy = cy + ((k & 2) ? anchor_radius : -anchor_radius);
            if (xx < 0 || yy < 0 || xx >= width || yy >= height)
                return fail(CHB_ERR_INVALID, "Image coordinate out of range: (%d, %d)", xx, yy);  // the reference panics here (src/shake.rs:325-328)
        }
    }
    chb_shake* sh = new chb_shake();
    sh->ctx = ctx;
    sh->W = width; sh->H = height; sh->C = channels;
    sh->n_anchors = n_anchors; sh->radius = anchor_radius; sh->search = search_radius;
    sh->size = 2 * anchor_radius + 1;
    sh->psize = 2 * (anchor_radius + search_radius) + 1;
    sh->search_size = 2 * search_radius + 1;
    sh->anchors.assign(anchors_xy, anchors_xy + 2 * n_anchors);
    const size_t win_bytes = (size_t)n_anchors * sh->size * sh->size * channels;
    const size_t patch_bytes = (size_t)n_anchors * sh->psize * sh->psize * channels;
    const size_t n_off = (size_t)sh->search_size * sh->search_size;
    auto bail = [&](cudaError_t e, const char* what) {
        chb_shake_destroy(sh);
        return fail(CHB_ERR_CUDA, "chb_shake_create: %s: %s", what, cudaGetErrorString(e));
    };
    cudaError_t e;
    if ((e = cudaSetDevice(ctx->devs[0].id)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaMalloc(&sh->d_windows, win_bytes)) != cudaSuccess) return bail(e, "windows");
    if ((e = cudaMalloc(&sh->d_patches, patch_bytes)) != cudaSuccess) return bail(e, "patches");
    if ((e = cudaMalloc(&sh->d_diffs, n_off * sizeof(int32_t))) != cudaSuccess) return bail(e, "diffs");
    if ((e = cudaMalloc(&sh->d_result, 2 * sizeof(int32_t))) != cudaSuccess) return bail(e, "result");
    if ((e = cudaMallocHost(&sh->h_patches, std::max(patch_bytes, win_bytes))) != cudaSuccess) return bail(e, "pinned patches");
    if ((e = cudaMallocHost(&sh->h_diffs, n_off * sizeof(int32_t))) != cudaSuccess) return bail(e, "pinned diffs");
    if ((e = cudaMallocHost(&sh->h_result, 2 * sizeof(int32_t))) != cudaSuccess) return bail(e, "pinned result");
    const size_t row = (size_t)sh->size * channels;
    for (int i = 0; i < n_anchors; i++) {
        const int x0 = anchors_xy[2 * i] - anchor_radius, y0 = anchors_xy[2 * i + 1] - anchor_radius;
        for (int dy = 0; dy < sh->size; dy++)
            memcpy(sh->h_patches + ((size_t)i * sh->size + dy) * row, first_frame + (size_t)(y0 + dy) * row_pitch + (size_t)x0 * channels, row);
    }
    if ((e = cudaMemcpyAsync(sh->d_windows, sh->h_patches, win_bytes, cudaMemcpyHostToDevice, sh->stream)) != cudaSuccess) return bail(e, "upload");
    if ((e = cudaStreamSynchronize(sh->stream)) != cudaSuccess) return bail(e, "sync");
    *out = sh;
    return CHB_OK;
}

// One frame of the par_iter in ShakeAnalyzer::analyze (src/shake.rs:248-283): calc_diffs over the search square and the
// first minimum. diffs (nullable) receives the (2s+1)^2 table. Thread-safe (calls on one analyzer are serialised).
extern "C" int chb_shake_offset(chb_shake* sh, const uint8_t* frame, size_t row_pitch, int32_t* out_dx, int32_t* out_dy, int32_t* diffs) {
    if (!sh || !frame || !out_dx || !out_dy) return fail(CHB_ERR_INVALID, "chb_shake_offset: null argument");
    if (row_pitch < (size_t)sh->W * sh->C) return fail(CHB_ERR_INVALID, "chb_shake_offset: row pitch smaller than a row");
    const int reach = sh->radius + sh->search;
    for (int i = 0; i < sh->n_anchors; i++) {
        const int cx = sh->anchors[2 * i], cy = sh->anchors[2 * i + 1];
        for (int k = 0; k < 4; k++) {
            const int xx = cx + ((k & 1) ? reach : -reach), yy = cy + ((k & 2) ? reach : -reach);
            if (xx < 0 || yy < 0 || xx >= sh->W || yy >= sh->H)
                return fail(CHB_ERR_INVALID, "Image coordinate out of range: (%d, %d)", xx, yy);  // src/shake.rs:371-374
        }
    }
    std::lock_guard<std::mutex> lk(sh->mu);
    CU(cudaSetDevice(sh->ctx->devs[0].id));
    const size_t row = (size_t)sh->psize * sh->C;
    for (int i = 0; i < sh->n_anchors; i++) {
        const int x0 = sh->anchors[2 * i] - reach, y0 = sh->anchors[2 * i + 1] - reach;
        for (int dy = 0; dy < sh->psize; dy++)
            memcpy(sh->h_patches + ((size_t)i * sh->psize + dy) * row, frame + (size_t)(y0 + dy) * row_pitch + (size_t)x0 * sh->C, row);
    }
    const size_t patch_bytes = (size_t)sh->n_anchors * sh->psize * row;
    const int n_off = sh->search_size * sh->search_size;
    CU(cudaMemcpyAsync(sh->d_patches, sh->h_patches, patch_bytes, cudaMemcpyHostToDevice, sh->stream));
    ShakeArgs a;
    a.windows = sh->d_windows; a.patches = sh->d_patches;
    a.n_anchors = sh->n_anchors; a.size = sh->size; a.psize = sh->psize; a.search_size = sh->search_size; a.C = sh->C;
    a.diffs = sh->d_diffs; a.result = sh->d_result;
    shake_diff_kernel<<<n_off, 256, 0, sh->stream>>>(a);
    shake_argmin_kernel<<<1, 256, 0, sh->stream>>>(a);
    g_launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sh->h_result, sh->d_result, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, sh->stream));
    if (diffs) CU(cudaMemcpyAsync(sh->h_diffs, sh->d_diffs, sizeof(int32_t) * (size_t)n_off, cudaMemcpyDeviceToHost, sh->stream));
    CU(cudaStreamSynchronize(sh->stream));
    if (diffs) memcpy(diffs, sh->h_diffs, sizeof(int32_t) * (size_t)n_off);
    const int min_idx = sh->h_result[0];
    *out_dx = (min_idx % sh->search_size) - sh->search;  // src/shake.rs:277-278
    *out_dy = (min_idx / sh->search_size) - sh->search;
    return CHB_OK;
}

extern "C" int chb_crop_create(const int32_t* off, int n, int width, int height, int32_t* out_xy, int32_t* out_w, int32_t* out_h) {
    // Crop::create, src/shake.rs:136-176
    int32_t xmin = 0, ymin = 0, xmax = 0, ymax = 0;
    for (int i = 0; i < n; i++) {
        xmin = std::min(xmin, off[2 * i]); xmax = std::max(xmax, off[2 * i]);
        ymin = std::min(ymin, off[2 * i + 1]); ymax = std::max(ymax, off[2 * i + 1]);
    }
    if (xmin == 0 && ymin == 0 && xmax == 0 && ymax == 0) return 0;
    *out_w = width + xmin - xmax;
    *out_h = height + ymin - ymax;
    for (int i = 0; i < n; i++) {
        out_xy[2 * i] = -xmin + off[2 * i];
        out_xy[2 * i + 1] = -ymin + off[2 * i + 1];
    }
    return 1;
}

extern "C" int chb_video_windows(int image_count, int in_has_start, int in_start, int in_has_end, int in_end, int in_step, int out_has_start,
                                 int out_start, int out_has_end, int out_end, int out_step, int32_t* win_start, int32_t* win_end,
                                 int32_t* out_number, int cap) {
    // create_video / create_video_simple, src/main.rs:230-286 and :349-404. Rust's % keeps the dividend's sign, like C++.
    if (in_step < 1 || out_step < 1) return -1;
    const int v_lower = out_has_start ? out_start : ((in_has_start && in_has_end) ? -(in_end - in_start) + 1 : 0);
    const int v_upper = out_has_end ? out_end : image_count;
    const int count = (v_upper - v_lower) / out_step;
    for (int i = 0; i < count && i < cap; i++) {
        const int frame = i * out_step + v_lower;
        int start = 0, end = image_count;
        if (in_has_start) {
            int st = frame + in_start;
            while (st < 0) st += in_step;
            start = std::max(st % in_step, frame + in_start);
        }
        if (in_has_end) end = std::min(image_count + (frame + in_end) % in_step - in_step, frame + in_end);
        win_start[i] = start;
        win_end[i] = end;
        out_number[i] = frame - v_lower;
    }
    return count < 0 ? 0 : count;
}
