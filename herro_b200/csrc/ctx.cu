// ctx.cu — hb_ctx: the C ABI of include/herro_b200.h on top of the kernels in features.cu /
// forward.cu.  Host-side responsibilities: replicate the read store, stage target batches
// (cross-read batching: the reference launches one tiny forward per read, src/features.rs:582,
// SURVEY.md F7), drive the kernel sequence on a stream, and re-assemble per-read segments
// the way consensus() does (src/consensus.rs:86-111,222-226).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include <pthread.h>
#include <sched.h>
#include <sys/stat.h>

#include "../../include/herro_b200.h"
#include "common.cuh"
#include "forward.h"
#include "windowing.h"
#include "torchscript.h"

using namespace hb;

namespace {

thread_local std::string g_create_err;
thread_local std::string* t_err_sink = nullptr;
std::atomic<uint64_t> g_ctx_generation{1};  // the launch worker reports into its own string

// allocation accounting (hb_stats.host_allocs / ms_host_alloc): page-locked and device allocations are slow and
// serialise with every other CUDA call of the process, so the steady state must not make any
std::atomic<uint64_t> g_allocs{0}, g_alloc_ns{0}, g_submit_wait_ns{0};
struct AllocScope {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const char* kind;
    size_t bytes;
    AllocScope(const char* k, size_t b) : kind(k), bytes(b) {}
    ~AllocScope() {
        const uint64_t ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        g_allocs.fetch_add(1, std::memory_order_relaxed);
        g_alloc_ns.fetch_add(ns, std::memory_order_relaxed);
        static const bool dbg = getenv("HERRO_B200_DEBUG_ALLOC") != nullptr;
        if (dbg) fprintf(stderr, "[herro_b200 alloc] %s %.1f MB %.2f ms\n", kind, (double)bytes / 1e6, (double)ns * 1e-6);
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;  // lane buffers: grown in stream order from the device's memory pool, because
                                // cudaMalloc/cudaFree synchronise the whole device and would stall the other lanes
    cudaError_t ensure(size_t bytes, bool keep = false) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = bytes + bytes / 2 + 256;
        AllocScope as_(st ? "device(stream-ordered)" : "device", ncap);
        void* np = nullptr;
        if (st) {
            cudaError_t e = cudaMallocAsync(&np, ncap, st);
            if (e != cudaSuccess) return e;
            if (keep && p && cap) cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
            if (p) cudaFreeAsync(p, st);
        } else {
            cudaError_t e = cudaMalloc(&np, ncap);
            if (e != cudaSuccess) return e;
            if (keep && p && cap) cudaMemcpy(np, p, cap, cudaMemcpyDeviceToDevice);
            if (p) cudaFree(p);
        }
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
    // grow to exactly `ncap` bytes without headroom (pre-sizing an idle lane from another lane's sizes); contents are kept
    // (the lane's last launch stays replayable / inspectable through the debug taps)
    cudaError_t reserve_exact(size_t ncap) {
        if (ncap <= cap) return cudaSuccess;
        AllocScope as_(st ? "device(stream-ordered, presize)" : "device(presize)", ncap);
        void* np = nullptr;
        cudaError_t e = st ? cudaMallocAsync(&np, ncap, st) : cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return e;
        if (p && cap) { if (st) cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st); else cudaMemcpy(np, p, cap, cudaMemcpyDeviceToDevice); }
        if (p) { if (st) cudaFreeAsync(p, st); else cudaFree(p); }
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        AllocScope as_("pinned(lane)", bytes + bytes / 2 + 256);
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t ncap = bytes + bytes / 2 + 256;
        cudaError_t e = cudaMallocHost(&p, ncap);
        if (e == cudaSuccess) cap = ncap;
        return e;
    }
    cudaError_t reserve_exact(size_t ncap) {
        if (ncap <= cap) return cudaSuccess;
        AllocScope as_("pinned(lane, presize)", ncap);
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, ncap);
        if (e == cudaSuccess) cap = ncap;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// grow-only pinned host array: target batches are staged directly in page-locked memory so the
// launch worker can cudaMemcpyAsync from them without a second copy
template <class T>
struct PinVec {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    int dev = 0;  // device whose context owns the allocation (set before first growth)
    PinVec() {}
    PinVec(const PinVec&) = delete;
    PinVec& operator=(const PinVec&) = delete;
    PinVec(PinVec&& o) noexcept : p(o.p), n(o.n), cap(o.cap), dev(o.dev) { o.p = nullptr; o.n = o.cap = 0; }
    PinVec& operator=(PinVec&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; dev = o.dev; o.p = nullptr; o.n = o.cap = 0; }
        return *this;
    }
    ~PinVec() { release(); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = cap = 0; }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        if (want * sizeof(T) > ((size_t)24 << 30)) return false;  // a staging array of > 24 GB is a sizing bug, not a workload: fail, do not pin
        size_t ncap = (n == 0) ? std::max<size_t>(want, 4096) : std::max<size_t>(want * 2, 4096);
        if (ncap * sizeof(T) > ((size_t)24 << 30)) ncap = want;
        AllocScope as_("pinned(staging)", ncap * sizeof(T));
        T* np = nullptr;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != dev) cudaSetDevice(dev);  // growth is rare: only then touch the runtime
        if (cudaHostAlloc((void**)&np, ncap * sizeof(T), cudaHostAllocPortable) != cudaSuccess) return false;
        if (n) memcpy(np, p, n * sizeof(T));
        if (p) cudaFreeHost(p);
        p = np;
        cap = ncap;
        return true;
    }
    bool push_back(const T& v) { if (!reserve(n + 1)) return false; p[n++] = v; return true; }
    bool append(const T* src, size_t k) { if (!reserve(n + k)) return false; if (k) memcpy(p + n, src, k * sizeof(T)); n += k; return true; }
    bool resize(size_t k) { if (!reserve(k)) return false; n = k; return true; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T* data() { return p; }
    const T* data() const { return p; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    void clear() { n = 0; }
    const T* begin() const { return p; }
    const T* end() const { return p + n; }
};

struct HostBatch {
    PinVec<DevTarget> tgt;
    PinVec<DevWin> win;
    PinVec<DevOverlap> ovl;
    PinVec<DevOW> ow;
    PinVec<uint8_t> cig;
    uint64_t op_cap = 0;      // op slots of host-windowed overlap-windows (assigned here)
    uint64_t raw_cap = 0;     // raw-op slots of device-windowed alignments (cig_len / 2 + 1 each)
    uint64_t dev_op_cap = 0;  // bound on the op slots their overlap-windows need (assigned by a scan on the device)
    uint32_t n_raw = 0;       // device-windowed alignments
    HostBatch() {}
    explicit HostBatch(int dev) { tgt.dev = win.dev = ovl.dev = ow.dev = cig.dev = dev; }
    void clear() { tgt.clear(); win.clear(); ovl.clear(); ow.clear(); cig.clear(); op_cap = raw_cap = dev_op_cap = 0; n_raw = 0; }
};

struct Result {
    uint32_t rid;
    int status;
    std::vector<uint32_t> seg_len;
    std::vector<uint8_t> seq;
    std::string msg;  // why status != HB_OK
};

struct LastLaunch {  // host copies of per-window metadata of the most recent launch (debug taps / replay)
    bool valid = false;
    std::vector<DevWin> win;
    std::vector<uint32_t> w_L, w_nsel, w_nsup;
    std::vector<uint64_t> w_rowbase, w_supbase;
    std::vector<uint32_t> ow_qid;  // KEEP_DEBUG: query read of every overlap-window (feature dump)
    std::unordered_map<uint64_t, uint32_t> index;  // (rid << 32 | wid) -> window
    uint64_t n_sup = 0, total_rows = 0;
    BatchView view;
};

}  // namespace

struct hb_ctx {
    int device = 0;
    hb_options opt{};
    std::string err;
    std::mutex mu;

    // weights
    FwdWeights wt{};
    std::vector<void*> weight_allocs;

    // read store
    bool have_reads = false;
    uint32_t n_reads = 0;
    std::vector<uint32_t> read_len;
    DevBuf d_words, d_word_off, d_len, d_qual, d_qual_off, d_ln;
    uint32_t ln_n = 0;
    ReadStoreView rs{};

    // staging: every submitting (feature) thread fills its own batch without taking the context lock;
    // full batches are handed to the launch worker's queue
    struct ThreadSlot { std::thread::id owner; HostBatch batch; uint32_t handed = 0; /* batches handed over since the last flush */ };
    std::vector<std::unique_ptr<ThreadSlot>> slots;
    PinBuf pin_in;  // staging of hb_upload_reads
    // Launch lanes (stream + device scratch + pinned result buffers each), one worker thread per lane:
    // while lane A's worker does its host work (copy-back, per-read reassembly), lane B's batch keeps the GPU busy.
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev[8]{};
        KTimer kt;
        PinBuf pin_small, pin_out;
        DevBuf d_tgt, d_win, d_ovl, d_ow, d_cig;
        DevBuf d_op_kl, d_op_t, d_op_q, d_ow_nops, d_ow_flags, d_ow_acc, d_ow_tend, d_col_ow, d_w_n1, d_w_S;
        DevBuf d_ovl_n, d_ovl_tot, d_ovl_score, d_sel_ow, d_w_nsel, d_rowmap, d_w_L, d_w_rowbase, d_w_nsup, d_w_reflmax;
        DevBuf d_mat_b, d_mat_q, d_row_emit, d_sup_row, d_sup_pk, d_w_supbase, d_fwd_win, d_fwd_row;
        DevBuf d_w_outlen, d_w_outoff, d_out, d_tgt_err, d_counters, d_ws, d_logits, d_info, d_big_key, d_big_cand, d_big_score;
        DevBuf d_raw_kl, d_raw_t, d_raw_q, d_aln_nops, d_aln_flags, d_ow_opoff;  // device windowing (windowing_dev.cu)
        DevBuf d_rank_ow;
        uint64_t rows_cap = 0;
        uint64_t seen_sizes = 0;  // version of hb_ctx::lane_sizes this lane has been pre-sized to
        LastLaunch last;
        std::thread worker;
        static constexpr int N_DEV = 51;
        void all_bufs(DevBuf* (&out)[N_DEV]) {
            DevBuf* bufs[N_DEV] = {&d_tgt, &d_win, &d_ovl, &d_ow, &d_cig, &d_op_kl, &d_op_t, &d_op_q, &d_ow_nops, &d_ow_flags, &d_ow_acc,
                                   &d_ow_tend, &d_col_ow, &d_w_n1, &d_w_S, &d_ovl_n, &d_ovl_tot, &d_ovl_score, &d_sel_ow, &d_w_nsel,
                                   &d_rowmap, &d_w_L, &d_w_rowbase, &d_w_nsup, &d_w_reflmax, &d_mat_b, &d_mat_q, &d_row_emit, &d_sup_row,
                                   &d_sup_pk, &d_w_supbase, &d_fwd_win, &d_fwd_row, &d_w_outlen, &d_w_outoff, &d_out, &d_tgt_err,
                                   &d_counters, &d_ws, &d_logits, &d_info, &d_big_key, &d_big_cand, &d_big_score,
                                   &d_raw_kl, &d_raw_t, &d_raw_q, &d_aln_nops, &d_aln_flags, &d_ow_opoff, &d_rank_ow};
            for (int i = 0; i < N_DEV; i++) out[i] = bufs[i];
        }
        void bind_stream() { DevBuf* bufs[N_DEV]; all_bufs(bufs); for (DevBuf* b : bufs) b->st = stream; }
        void release() {
            DevBuf* bufs[N_DEV]; all_bufs(bufs);
            for (DevBuf* b : bufs) b->release();
            pin_small.release(); pin_out.release();
            for (auto& e : ev) if (e) cudaEventDestroy(e);
            kt.destroy();
            if (stream) cudaStreamDestroy(stream);
        }
    };
    // Largest buffer capacities any lane has needed so far.  A lane that has not run yet (or ran smaller batches) grows its
    // buffers to these while it is idle, so that its first real batch allocates nothing (r01: 26 allocations / 226 ms inside the
    // timed region at 8 GPUs, when host contention made the third lane start its first batch there).
    struct LaneSizes { size_t dev[Lane::N_DEV] = {0}; size_t pin_small = 0, pin_out = 0; uint64_t rows_cap = 0; };
    LaneSizes lane_sizes;
    uint64_t lane_sizes_version = 0;
    static constexpr int MAX_LANES = 4;
    Lane lanes[MAX_LANES];
    int n_lanes = 3;     // HERRO_B200_LANES overrides (1..4)
    uint32_t min_launch = 256;  // smallest per-thread hand-over unless launch_targets itself is smaller (HERRO_B200_MIN_LAUNCH: experiments)
    int last_lane = -1;  // lane of the most recently finished launch (debug taps / replay)
    uint32_t chunk_pos = 65536;  // supported positions per forward pass (HERRO_B200_CHUNK_POS): one pass per launch unless huge

    std::deque<Result> results;
    hb_stats stats{};

    // launch worker: batches are processed asynchronously so that the host can stage batch i+1
    // (and drain results of batch i-1) while batch i is on the GPU
    std::condition_variable cv_work, cv_idle;
    std::deque<HostBatch> queue;
    std::vector<HostBatch> pool;  // recycled staging batches (keep their pinned capacity)
    bool stop = false;
    int busy = 0;  // lanes currently inside a launch
    int worker_rc = HB_OK;
    std::string worker_err;
    bool idle() const { return queue.empty() && busy == 0; }
    size_t cap_hint[5] = {0, 0, 0, 0, 0};  // largest batch array sizes seen (tgt, win, ovl, ow, cig)
    std::atomic<uint32_t> handed_total{0};  // batches handed over by all threads since the last flush (slow-start ramp)
    std::atomic<uint32_t> n_slots{0};       // submitting threads registered since the last flush
    std::atomic<bool> time_kernels{false};  // hb_set_kernel_timing
    uint64_t alloc_base[3] = {0, 0, 0};     // g_allocs / g_alloc_ns / g_submit_wait_ns at the last hb_reset_stats
    uint64_t generation = 0;  // distinguishes contexts that reuse an address (thread-local slot cache)
    // debugging aids read from the environment once, in hb_create
    bool host_windowing = false;     // HERRO_B200_HOST_WINDOWING: hb_submit_alignments runs extract_windows on the host (A-B test)
    bool pileup_v1 = false;          // HERRO_B200_PILEUP_V1: the former position-walk pileup kernel (A-B parity test)
    uint32_t arena_rows_per_win = 0; // HERRO_B200_ARENA_ROWS: initial row-arena rows per window (default 1.5 W); tests shrink it
                                     // to force the overflow -> regrow -> relaunch path
    // staging-batch pool: every HostBatch that exists is counted, so the steady state allocates nothing
    uint32_t batches_alive = 0;
    // CPUs of the NUMA node the GPU hangs off (empty: unknown); launch workers are bound to them, hb_bind_calling_thread
    // does the same for the host's feature / consumer threads
    cpu_set_t node_cpus;
    bool have_node_cpus = false;
    int numa_node = -1;
};

namespace {

std::string& errref(hb_ctx* ctx) { return t_err_sink ? *t_err_sink : ctx->err; }

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            errref(ctx) = std::string(#call) + ": " + cudaGetErrorString(e__);        \
            return HB_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

int fail(hb_ctx* ctx, int code, const std::string& msg) {
    errref(ctx) = msg;
    return code;
}

// ---------------------------------------------------------------------------------- weights
#pragma pack(push, 1)
struct BlobHeader {
    char magic[8];
    uint32_t version, n_tensors;
    uint32_t cfg[16];
};
struct BlobEntry {
    char name[48];
    uint32_t dtype, ndim, shape[4];
    uint64_t offset, nbytes;
};
#pragma pack(pop)

// NUMA node of the GPU (sysfs) and that node's CPU list: ranks sharing a box must not pile their host threads onto one socket
void probe_numa(hb_ctx* ctx) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, ctx->device) != cudaSuccess) return;
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0) return;
    path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    f = fopen(path.c_str(), "r");
    if (!f) return;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!got) return;
    cpu_set_t allowed, set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
    int n = 0;
    for (const char* p = line; *p;) {  // "0-31,64-95"
        char* e;
        long a = strtol(p, &e, 10), b2 = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b2 = strtol(p, &e, 10); }
        for (long c = a; c <= b2 && c < CPU_SETSIZE; c++)
            if (CPU_ISSET((int)c, &allowed)) { CPU_SET((int)c, &set); n++; }
        p = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    if (n == 0) return;
    ctx->node_cpus = set;
    ctx->have_node_cpus = true;
    ctx->numa_node = node;
}

// A model file as the canonical tensor table (names and forms of herro_b200/weights.py): either the HB200W1 blob, or a
// TorchScript archive of the same architecture - what the reference's `-m` names (src/inference.rs:185) - read by torchscript.cpp.
struct ModelFile {
    uint32_t cfg[10] = {0};  // tokens, emb, reads, stem_k, C, H, layers, F, D, classes
    std::map<std::string, std::vector<float>> T;
};
int read_model_file(const char* path, ModelFile& mf, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open model file ") + path; return HB_ERR_MODEL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); err = "cannot size model file"; return HB_ERR_MODEL; }
    std::vector<uint8_t> buf((size_t)sz);
    if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); err = "short read on model file"; return HB_ERR_MODEL; }
    fclose(f);
    if (ts_is_zip(buf.data(), buf.size())) {
        TsModel m;
        if (!ts_read_archive(buf.data(), buf.size(), m)) { err = "TorchScript archive: " + m.err; return HB_ERR_MODEL; }
        TsDims d;
        if (!ts_to_canonical(m, 4, d, mf.T, err)) { err = "TorchScript archive: " + err; return HB_ERR_MODEL; }
        const uint32_t c[10] = {12, 6, 31, (uint32_t)d.stem_k, (uint32_t)d.channels, (uint32_t)d.heads, (uint32_t)d.layers, (uint32_t)d.ffn, (uint32_t)d.collapse, 5};
        memcpy(mf.cfg, c, sizeof c);
        return HB_OK;
    }
    if ((size_t)sz < sizeof(BlobHeader)) { err = "model file too small"; return HB_ERR_MODEL; }
    BlobHeader h;
    memcpy(&h, buf.data(), sizeof h);
    if (memcmp(h.magic, "HB200W1\0", 8) != 0 || h.version != 1) {
        err = "neither an HB200W1 weights blob nor a TorchScript archive";
        return HB_ERR_MODEL;
    }
    memcpy(mf.cfg, h.cfg, sizeof mf.cfg);
    for (uint32_t i = 0; i < h.n_tensors; i++) {
        BlobEntry e;
        size_t eo = sizeof(BlobHeader) + (size_t)i * sizeof(BlobEntry);
        if (eo + sizeof e > (size_t)sz) { err = "truncated tensor table"; return HB_ERR_MODEL; }
        memcpy(&e, buf.data() + eo, sizeof e);
        if (e.dtype != 0 || e.offset > (uint64_t)sz || e.nbytes > (uint64_t)sz - e.offset || (e.offset & 3u) || (e.nbytes & 3u)) {
            err = "bad tensor entry";
            return HB_ERR_MODEL;
        }
        char nm[49];
        memcpy(nm, e.name, 48);
        nm[48] = 0;
        std::vector<float>& v = mf.T[nm];
        v.resize((size_t)(e.nbytes / 4));
        memcpy(v.data(), buf.data() + e.offset, (size_t)e.nbytes);
    }
    return HB_OK;
}

int load_weights(hb_ctx* ctx, const char* path) {
    ModelFile mf;
    {
        std::string e;
        const int rc = read_model_file(path, mf, e);
        if (rc != HB_OK) return fail(ctx, rc, e);
    }
    const uint32_t* cfg = mf.cfg;
    if (cfg[0] != 12 || cfg[1] != 6 || cfg[2] != 31 || cfg[9] != 5)
        return fail(ctx, HB_ERR_MODEL, "unsupported fixed dimensions (tokens/emb/reads/classes)");
    FwdWeights& wt = ctx->wt;
    wt.stem_k = (int)cfg[3]; wt.C = (int)cfg[4]; wt.H = (int)cfg[5]; wt.layers = (int)cfg[6];
    wt.F = (int)cfg[7]; wt.D = (int)cfg[8];
    if (wt.layers < 1 || wt.layers > MAX_LAYERS || wt.H < 1 || wt.C % wt.H || (wt.C / wt.H != 16 && wt.C / wt.H != 32) ||
        wt.C % 128 || wt.F % 128 || wt.D % 128 || !(wt.stem_k & 1) || wt.stem_k > 129 || wt.C > 1024)
        return fail(ctx, HB_ERR_MODEL, "unsupported model dimensions (need C,F,D % 128 == 0, head_dim 16 or 32, odd stem_k)");
    std::unordered_map<std::string, std::pair<const float*, size_t>> T;
    for (auto& kv : mf.T) T[kv.first] = {kv.second.data(), kv.second.size()};
    auto need = [&](const std::string& n, size_t count, const float*& hostp) -> bool {
        auto it = T.find(n);
        if (it == T.end() || it->second.second != count) { ctx->err = "missing/mis-sized tensor " + n; return false; }
        hostp = it->second.first;
        return true;
    };
    auto upload = [&](const float* hostp, size_t count, const float*& devp) -> bool {
        void* d = nullptr;
        if (cudaMalloc(&d, count * 4) != cudaSuccess) { ctx->err = "cudaMalloc(weights)"; return false; }
        ctx->weight_allocs.push_back(d);
        if (cudaMemcpy(d, hostp, count * 4, cudaMemcpyHostToDevice) != cudaSuccess) { ctx->err = "cudaMemcpy(weights)"; return false; }
        devp = (const float*)d;
        return true;
    };
    auto get = [&](const std::string& n, size_t count, const float*& devp) -> bool {
        const float* hp;
        return need(n, count, hp) && upload(hp, count, devp);
    };
    const int C = wt.C, K = wt.stem_k, F = wt.F, D = wt.D;
    const float *emb, *stem_w;
    if (!need("emb", 12 * 6, emb) || !need("stem_w", (size_t)C * 7 * K, stem_w)) return HB_ERR_MODEL;
    // fold the embedding into the conv: tab[j][t][c] = sum_e stem_w[c][e][j] * emb[t][e]
    std::vector<float> tab((size_t)K * 12 * C), wq((size_t)K * C);
    for (int j = 0; j < K; j++)
        for (int c = 0; c < C; c++) {
            for (int t = 0; t < 12; t++) {
                float acc = 0.f;  // fp32, e ascending: the order a direct conv over 7 input channels would use
                for (int e = 0; e < 6; e++) acc = fmaf(stem_w[((size_t)c * 7 + e) * K + j], emb[t * 6 + e], acc);
                tab[((size_t)j * 12 + t) * C + c] = acc;
            }
            wq[(size_t)j * C + c] = stem_w[((size_t)c * 7 + 6) * K + j];
        }
    if (!upload(tab.data(), tab.size(), wt.stem_tab) || !upload(wq.data(), wq.size(), wt.stem_wq)) return HB_ERR_CUDA;
    bool ok = get("stem_b", C, wt.stem_b) && get("read_pos", 31 * (size_t)C, wt.read_pos);
    for (int l = 0; ok && l < wt.layers; l++) {
        const std::string p = "l" + std::to_string(l) + ".";
        FwdLayer& ly = wt.layer[l];
        ok = get(p + "ln1_g", C, ly.ln1_g) && get(p + "ln1_b", C, ly.ln1_b) && get(p + "wqkv", (size_t)3 * C * C, ly.wqkv) &&
             get(p + "bqkv", 3 * (size_t)C, ly.bqkv) && get(p + "wo", (size_t)C * C, ly.wo) && get(p + "bo", C, ly.bo) &&
             get(p + "ln2_g", C, ly.ln2_g) && get(p + "ln2_b", C, ly.ln2_b) && get(p + "w1", (size_t)F * C, ly.w1) &&
             get(p + "b1", F, ly.b1) && get(p + "w2", (size_t)C * F, ly.w2) && get(p + "b2", C, ly.b2);
    }
    ok = ok && get("lnf_g", C, wt.lnf_g) && get("lnf_b", C, wt.lnf_b) && get("wc", (size_t)D * 31 * C, wt.wc) &&
         get("bc", D, wt.bc) && get("wb", 5 * (size_t)D, wt.wb) && get("bb", 5, wt.bb) && get("wi", D, wt.wi) &&
         get("bi", 1, wt.bi);
    if (!ok) return ctx->err.rfind("cuda", 0) == 0 ? HB_ERR_CUDA : HB_ERR_MODEL;
    // bf16 hi/lo split of the contraction weights for the tcgen05 path (gemm_tc.cu)
    if (cudaDeviceGetAttribute(&wt.num_sms, cudaDevAttrMultiProcessorCount, ctx->device) != cudaSuccess) wt.num_sms = 148;
    auto split = [&](const float* w, size_t n, SplitW& s) -> bool {
        void *hi = nullptr, *lo = nullptr;
        if (split_weights(w, n, &hi, &lo) != cudaSuccess) { ctx->err = "cuda: weight split failed"; return false; }
        ctx->weight_allocs.push_back(hi);
        ctx->weight_allocs.push_back(lo);
        s.hi = hi; s.lo = lo;
        return true;
    };
    for (int l = 0; l < wt.layers; l++) {
        FwdLayer& ly = wt.layer[l];
        if (!split(ly.wqkv, (size_t)3 * C * C, ly.s_qkv) || !split(ly.wo, (size_t)C * C, ly.s_o) ||
            !split(ly.w1, (size_t)F * C, ly.s_1) || !split(ly.w2, (size_t)C * F, ly.s_2)) return HB_ERR_CUDA;
    }
    if (!split(wt.wc, (size_t)D * 31 * C, wt.s_c)) return HB_ERR_CUDA;
    // head-grouped copy of Wqkv / bqkv for the fused QKV+attention kernel: row (h, s, d) = row s*C + h*32 + d (s = q,k,v)
    if (C == 128 && wt.H == 4) {
        for (int l = 0; l < wt.layers; l++) {
            const std::string p = "l" + std::to_string(l) + ".";
            const float *hw, *hbias;
            if (!need(p + "wqkv", (size_t)3 * C * C, hw) || !need(p + "bqkv", 3 * (size_t)C, hbias)) return HB_ERR_MODEL;
            std::vector<float> wp((size_t)3 * C * C), bp((size_t)3 * C);
            for (int h = 0; h < 4; h++)
                for (int sI = 0; sI < 3; sI++)
                    for (int d = 0; d < 32; d++) {
                        const size_t dst = (size_t)h * 96 + sI * 32 + d, src = (size_t)sI * C + h * 32 + d;
                        memcpy(&wp[dst * C], hw + src * C, (size_t)C * 4);
                        bp[dst] = hbias[src];
                    }
            const float* dwp = nullptr;
            if (!upload(wp.data(), wp.size(), dwp) || !upload(bp.data(), bp.size(), wt.layer[l].bqkvp)) return HB_ERR_CUDA;
            if (!split(dwp, wp.size(), wt.layer[l].s_qkvp)) return HB_ERR_CUDA;
        }
    }
    // W' of the tensor-core stem: [C][taps*16] = tab (11 token slots), wq twice (q_hi, q_lo columns), zero padding
    wt.stem_kblocks = 0;
    if (C == 128 && K <= 64 && !getenv("HERRO_B200_STEM_SIMT")) {
        const int kbl = (K * 16 + 63) / 64, Kp = kbl * 64;
        std::vector<float> wp((size_t)C * Kp, 0.f);
        for (int c = 0; c < C; c++)
            for (int j = 0; j < K; j++) {
                for (int t = 0; t < 11; t++) wp[(size_t)c * Kp + j * 16 + t] = tab[((size_t)j * 12 + t) * C + c];
                wp[(size_t)c * Kp + j * 16 + 11] = wq[(size_t)j * C + c];
                wp[(size_t)c * Kp + j * 16 + 12] = wq[(size_t)j * C + c];
            }
        const float* dwp = nullptr;
        if (!upload(wp.data(), wp.size(), dwp)) return HB_ERR_CUDA;
        if (!split(dwp, wp.size(), wt.s_stem)) return HB_ERR_CUDA;
        wt.stem_kblocks = kbl;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { ctx->err = "cuda: weight split kernel failed"; return HB_ERR_CUDA; }
    return HB_OK;
}

// ---------------------------------------------------------------------------------- batch run
template <class T>
size_t vbytes(const PinVec<T>& v) { return v.size() * sizeof(T); }

int ensure_batch_buffers(hb_ctx* ctx, hb_ctx::Lane* L, const HostBatch& hbt) {
    const size_t nt = hbt.tgt.size(), nw = hbt.win.size(), no = hbt.ovl.size(), now_ = hbt.ow.size();
    const uint32_t W = ctx->opt.window_size;
    CK(L->d_tgt.ensure(nt * sizeof(DevTarget)));
    CK(L->d_win.ensure(nw * sizeof(DevWin)));
    CK(L->d_ovl.ensure(std::max<size_t>(no, 1) * sizeof(DevOverlap)));
    CK(L->d_ow.ensure(std::max<size_t>(now_, 1) * sizeof(DevOW)));
    CK(L->d_cig.ensure(std::max<size_t>(hbt.cig.size(), 16)));
    const size_t opc = std::max<uint64_t>(hbt.op_cap + hbt.dev_op_cap, 1);
    if (opc >= 0xffffffffull) return fail(ctx, HB_ERR_CAPACITY, "batch too large: more than 2^32 CIGAR ops (lower launch_targets)");
    const size_t rawc = std::max<uint64_t>(hbt.raw_cap, 1);
    CK(L->d_raw_kl.ensure(rawc * 4));
    CK(L->d_raw_t.ensure(rawc * 4));
    CK(L->d_raw_q.ensure(rawc * 4));
    CK(L->d_op_kl.ensure(opc * 4));
    CK(L->d_op_t.ensure(opc * 4));
    CK(L->d_op_q.ensure(opc * 4));
    const size_t ow1 = std::max<size_t>(now_, 1);
    CK(L->d_ow_nops.ensure(ow1 * 4));
    CK(L->d_ow_flags.ensure(ow1 * 4));
    CK(L->d_ow_acc.ensure(ow1 * 4));
    CK(L->d_ow_tend.ensure(ow1 * 4));
    CK(L->d_col_ow.ensure(ow1 * 4));
    CK(L->d_big_key.ensure(ow1 * 4));
    CK(L->d_big_cand.ensure(ow1 * 4));
    CK(L->d_big_score.ensure(ow1 * 8));
    CK(L->d_ow_opoff.ensure(ow1 * 8));
    CK(L->d_rank_ow.ensure(ow1 * 4));
    CK(L->d_w_n1.ensure(nw * 4));
    CK(L->d_w_S.ensure(nw * 4));
    const size_t no1 = std::max<size_t>(no, 1);
    CK(L->d_ovl_n.ensure(no1 * 4));
    CK(L->d_ovl_tot.ensure(no1 * 4));
    CK(L->d_ovl_score.ensure(no1 * 8));
    CK(L->d_aln_nops.ensure(no1 * 4));
    CK(L->d_aln_flags.ensure(no1 * 4));
    CK(L->d_sel_ow.ensure(nw * TOP_K * 4));
    CK(L->d_w_nsel.ensure(nw * 4));
    CK(L->d_rowmap.ensure(nw * (size_t)(W + 1) * 4));
    CK(L->d_w_L.ensure(nw * 4));
    CK(L->d_w_rowbase.ensure(nw * 8));
    CK(L->d_w_nsup.ensure(nw * 4));
    CK(L->d_w_reflmax.ensure(nw * 4));
    CK(L->d_w_supbase.ensure(nw * 8));
    CK(L->d_w_outlen.ensure(nw * 4));
    CK(L->d_w_outoff.ensure(nw * 8));
    CK(L->d_tgt_err.ensure(nt * 4));
    CK(L->d_counters.ensure(CNT_N * 4));
    return HB_OK;
}

int ensure_row_buffers(hb_ctx* ctx, hb_ctx::Lane* L, uint64_t rows) {
    if (rows <= L->rows_cap) return HB_OK;
    CK(L->d_mat_b.ensure(rows * ROW_BYTES));
    CK(L->d_mat_q.ensure(rows * ROW_BYTES));
    CK(L->d_row_emit.ensure(rows));
    CK(L->d_sup_row.ensure(rows * 4));
    CK(L->d_sup_pk.ensure(rows * 4));
    CK(L->d_out.ensure(rows));
    L->rows_cap = rows;
    return HB_OK;
}

// the pointer fields of a view: the lane's buffers as they are now
void set_view_ptrs(hb_ctx* ctx, hb_ctx::Lane* L, BatchView& b) {
    b.rs = ctx->rs;
    b.tgt = L->d_tgt.as<DevTarget>();
    b.win = L->d_win.as<DevWin>();
    b.ovl = L->d_ovl.as<DevOverlap>();
    b.ow = L->d_ow.as<DevOW>();
    b.ow_mut = L->d_ow.as<DevOW>();
    b.raw_kl = L->d_raw_kl.as<uint32_t>();
    b.raw_t = L->d_raw_t.as<uint32_t>();
    b.raw_q = L->d_raw_q.as<uint32_t>();
    b.aln_nops = L->d_aln_nops.as<uint32_t>();
    b.aln_flags = L->d_aln_flags.as<uint32_t>();
    b.ow_opoff = L->d_ow_opoff.as<uint64_t>();
    b.cig = L->d_cig.as<uint8_t>();
    b.op_kl = L->d_op_kl.as<uint32_t>();
    b.op_t = L->d_op_t.as<uint32_t>();
    b.op_q = L->d_op_q.as<uint32_t>();
    b.ow_nops = L->d_ow_nops.as<uint32_t>();
    b.ow_flags = L->d_ow_flags.as<uint32_t>();
    b.ow_acc = L->d_ow_acc.as<float>();
    b.ow_tend = L->d_ow_tend.as<uint32_t>();
    b.col_ow = L->d_col_ow.as<uint32_t>();
    b.big_key = L->d_big_key.as<float>();
    b.big_cand = L->d_big_cand.as<uint32_t>();
    b.big_score = L->d_big_score.as<double>();
    b.w_n1 = L->d_w_n1.as<uint32_t>();
    b.w_S = L->d_w_S.as<uint32_t>();
    b.ovl_n = L->d_ovl_n.as<uint32_t>();
    b.ovl_tot = L->d_ovl_tot.as<uint32_t>();
    b.ovl_score = L->d_ovl_score.as<double>();
    b.ln_table = ctx->d_ln.as<double>();
    b.ln_table_n = ctx->ln_n;
    b.rank_ow = L->d_rank_ow.as<uint32_t>();
    b.sel_ow = L->d_sel_ow.as<uint32_t>();
    b.w_nsel = L->d_w_nsel.as<uint32_t>();
    b.rowmap = L->d_rowmap.as<uint32_t>();
    b.w_L = L->d_w_L.as<uint32_t>();
    b.w_rowbase = L->d_w_rowbase.as<uint64_t>();
    b.w_nsup = L->d_w_nsup.as<uint32_t>();
    b.w_reflmax = L->d_w_reflmax.as<uint32_t>();
    b.mat_bases = L->d_mat_b.as<uint8_t>();
    b.mat_quals = L->d_mat_q.as<uint8_t>();
    b.row_emit = L->d_row_emit.as<uint8_t>();
    b.sup_row = L->d_sup_row.as<uint32_t>();
    b.sup_pk = L->d_sup_pk.as<uint32_t>();
    b.w_supbase = L->d_w_supbase.as<uint64_t>();
    b.fwd_win = L->d_fwd_win.as<uint32_t>();
    b.fwd_row = L->d_fwd_row.as<uint32_t>();
    b.w_outlen = L->d_w_outlen.as<uint32_t>();
    b.w_outoff = L->d_w_outoff.as<uint64_t>();
    b.out_bytes = L->d_out.as<uint8_t>();
    b.tgt_err = L->d_tgt_err.as<uint32_t>();
    b.counters = L->d_counters.as<uint32_t>();
}

BatchView make_view(hb_ctx* ctx, hb_ctx::Lane* L, const HostBatch& hbt) {
    BatchView b{};
    b.W = ctx->opt.window_size;
    b.n_tgt = (uint32_t)hbt.tgt.size();
    b.n_win = (uint32_t)hbt.win.size();
    b.n_ovl = (uint32_t)hbt.ovl.size();
    b.n_ow = (uint32_t)hbt.ow.size();
    b.batch_size = ctx->opt.batch_size;
    b.rows_cap = L->rows_cap;
    b.n_raw = hbt.n_raw;
    b.op_base_dev = (uint32_t)hbt.op_cap;
    set_view_ptrs(ctx, L, b);
    return b;
}

int zero_scratch(hb_ctx* ctx, hb_ctx::Lane* L, const BatchView& b) {
    CK(cudaMemsetAsync(b.ovl_n, 0, std::max<size_t>(b.n_ovl, 1) * 4, L->stream));
    CK(cudaMemsetAsync(b.ovl_tot, 0, std::max<size_t>(b.n_ovl, 1) * 4, L->stream));
    CK(cudaMemsetAsync(b.tgt_err, 0, (size_t)b.n_tgt * 4, L->stream));
    CK(cudaMemsetAsync(b.counters, 0, CNT_N * 4, L->stream));
    return HB_OK;
}

// The forward + consensus part once the number of supported positions is known.
int launch_tail(hb_ctx* ctx, hb_ctx::Lane* L, const BatchView& b, uint64_t n_sup, uint64_t* launches) {
    *launches += launch_features_c2(b, L->stream, L->kt);
    for (uint64_t n0 = 0; n0 < n_sup; n0 += ctx->chunk_pos) {
        const uint32_t np = (uint32_t)std::min<uint64_t>(ctx->chunk_pos, n_sup - n0);
        *launches += launch_forward_chunk(b, ctx->wt, (uint32_t)n0, np, L->d_ws.as<uint8_t>(), L->d_logits.as<float>(),
                                          L->d_info.as<float>(), L->stream, L->kt);
    }
    CK(cudaEventRecord(L->ev[4], L->stream));
    *launches += launch_consensus(b, L->stream, L->kt);
    CK(cudaEventRecord(L->ev[5], L->stream));
    return HB_OK;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int run_batch(hb_ctx* ctx, hb_ctx::Lane* L, HostBatch& hbt) {
    if (hbt.tgt.empty()) return HB_OK;
    const double t_begin = now_ms();
    double t_wait = 0, t_mark = t_begin;
    hb_stats S{};  // merged into ctx->stats under the lock at the end
#define SYNC_TIMED() do { const double t__ = now_ms(); CK(cudaStreamSynchronize(L->stream)); t_wait += now_ms() - t__; } while (0)
#define PHASE(i) do { const double t__ = now_ms(); S.ms_worker_phase[i] += t__ - t_mark; t_mark = t__; } while (0)
    std::vector<Result> out_results;
    const uint32_t W = ctx->opt.window_size;
    int rc = ensure_batch_buffers(ctx, L, hbt);
    if (rc) return rc;
    const size_t nt = hbt.tgt.size(), nw = hbt.win.size();
    // ---- H2D straight from the pinned staging arrays of the batch
    const size_t sz[5] = {vbytes(hbt.tgt), vbytes(hbt.win), vbytes(hbt.ovl), vbytes(hbt.ow), hbt.cig.size()};
    const void* src[5] = {hbt.tgt.data(), hbt.win.data(), hbt.ovl.data(), hbt.ow.data(), hbt.cig.data()};
    void* dst[5] = {L->d_tgt.p, L->d_win.p, L->d_ovl.p, L->d_ow.p, L->d_cig.p};
    for (int i = 0; i < 5; i++)
        if (sz[i]) CK(cudaMemcpyAsync(dst[i], src[i], sz[i], cudaMemcpyHostToDevice, L->stream));
    S.h2d_bytes += sz[0] + sz[1] + sz[2] + sz[3] + sz[4];

    if (L->rows_cap == 0) {
        const uint64_t per_win = ctx->arena_rows_per_win ? ctx->arena_rows_per_win : (uint64_t)W + W / 2;
        rc = ensure_row_buffers(ctx, L, (uint64_t)nw * per_win + 64);
        if (rc) return rc;
    }
    CK(L->pin_small.ensure(CNT_N * 4 + nw * 4 * 4 + nt * 4 + nw * TOP_K * 4 + 1024));
    uint32_t* h_cnt = L->pin_small.as<uint32_t>();
    uint64_t launches = 0;
    BatchView b;
    uint64_t total_rows = 0;
    PHASE(0);
    L->kt.on = ctx->time_kernels.load(std::memory_order_relaxed);
    L->kt.st = L->stream;
    for (int attempt = 0;; attempt++) {
        if (attempt) L->kt.discard();
        b = make_view(ctx, L, hbt);
        rc = zero_scratch(ctx, L, b);
        if (rc) return rc;
        CK(cudaEventRecord(L->ev[0], L->stream));
        launches += launch_features_a(b, L->stream, L->kt);
        CK(cudaEventRecord(L->ev[1], L->stream));
        launches += launch_pileup(b, L->stream, L->kt, ctx->pileup_v1);
        CK(cudaEventRecord(L->ev[2], L->stream));
        launches += launch_features_c1(b, L->stream, L->kt);  // ref_lmax + scan; the work list needs its buffers first
        CK(cudaMemcpyAsync(h_cnt, b.counters, CNT_N * 4, cudaMemcpyDeviceToHost, L->stream));
        PHASE(1);
        SYNC_TIMED();
        PHASE(2);
        total_rows = (uint64_t)h_cnt[CNT_TOTAL_ROWS] | ((uint64_t)h_cnt[CNT_TOTAL_ROWS + 1] << 32);
        if (!h_cnt[CNT_OVERFLOW]) break;
        if (attempt >= 2) return fail(ctx, HB_ERR_CAPACITY, "row arena overflow persisted after regrowth");
        rc = ensure_row_buffers(ctx, L, total_rows + total_rows / 8 + 4096);
        if (rc) return rc;
    }
    const uint64_t n_sup = (uint64_t)h_cnt[CNT_NSUP] | ((uint64_t)h_cnt[CNT_NSUP + 1] << 32);
    CK(L->d_fwd_win.ensure(std::max<uint64_t>(n_sup, 1) * 4));
    CK(L->d_fwd_row.ensure(std::max<uint64_t>(n_sup, 1) * 4));
    CK(L->d_logits.ensure(std::max<uint64_t>(n_sup, 1) * 5 * 4));
    CK(L->d_info.ensure(std::max<uint64_t>(n_sup, 1) * 4));
    CK(L->d_ws.ensure(fwd_workspace_bytes(ctx->wt, (uint32_t)std::min<uint64_t>(ctx->chunk_pos, std::max<uint64_t>(n_sup, 1)))));
    b = make_view(ctx, L, hbt);
    CK(cudaEventRecord(L->ev[3], L->stream));
    rc = launch_tail(ctx, L, b, n_sup, &launches);
    if (rc) return rc;

    // ---- D2H: per-window metadata, then exactly the emitted bytes
    uint32_t* h_outlen = h_cnt + CNT_N;
    uint32_t* h_nsel = h_outlen + nw;
    uint32_t* h_L = h_nsel + nw;
    uint32_t* h_nsup = h_L + nw;
    uint32_t* h_terr = h_nsup + nw;
    uint32_t* h_sel = h_terr + nt;
    CK(cudaMemcpyAsync(h_cnt, b.counters, CNT_N * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaMemcpyAsync(h_outlen, b.w_outlen, nw * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaMemcpyAsync(h_nsel, b.w_nsel, nw * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaMemcpyAsync(h_L, b.w_L, nw * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaMemcpyAsync(h_nsup, b.w_nsup, nw * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaMemcpyAsync(h_terr, b.tgt_err, nt * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaMemcpyAsync(h_sel, b.sel_ow, nw * TOP_K * 4, cudaMemcpyDeviceToHost, L->stream));
    // the emitted bytes of all windows are contiguous from offset 0 and number at most one per matrix row, so the
    // row count (known since the first wait) bounds the copy: no second round trip for the exact size
    CK(L->pin_out.ensure(total_rows + 16));
    if (total_rows) CK(cudaMemcpyAsync(L->pin_out.p, b.out_bytes, total_rows, cudaMemcpyDeviceToHost, L->stream));
    PHASE(3);
    SYNC_TIMED();
    PHASE(4);
    const uint64_t total_out = (uint64_t)h_cnt[CNT_TOTAL_OUT] | ((uint64_t)h_cnt[CNT_TOTAL_OUT + 1] << 32);
    if (total_out > total_rows) return fail(ctx, HB_ERR_CAPACITY, "consensus emitted more bytes than matrix rows");
    S.d2h_bytes += CNT_N * 4 + nw * 16 + nt * 4 + nw * TOP_K * 4 + total_rows;

    // ---- timing
    float ms;
    cudaEventElapsedTime(&ms, L->ev[0], L->ev[2]); S.ms_features += ms;
    cudaEventElapsedTime(&ms, L->ev[3], L->ev[4]); S.ms_forward += ms;
    cudaEventElapsedTime(&ms, L->ev[4], L->ev[5]); S.ms_consensus += ms;
    L->kt.collect(S.ms_kernel, S.n_kernel);
    L->kt.on = false;
    {
        uint64_t gf = 0;
        const uint64_t ff = forward_flops_per_pos(ctx->wt, &gf);
        S.forward_flops += ff * n_sup;
        S.gemm_flops += gf * n_sup;
        uint64_t cf[16];
        forward_class_flops_per_pos(ctx->wt, cf);
        for (int i = 0; i < HB_NUM_KERNEL_CLASSES; i++) S.class_flops[i] += cf[i] * n_sup;
    }

    // ---- per-read reassembly (src/consensus.rs:90-111,222-226)
    const uint8_t* outb = L->pin_out.as<uint8_t>();
    uint64_t o = 0, corrected = 0, algo = 0;
    for (size_t t = 0; t < nt; t++) {
        const DevTarget& tg = hbt.tgt[t];
        Result r;
        r.rid = tg.rid;
        r.status = HB_OK;
        if (h_terr[t] & TERR_BAD_INPUT) {
            r.status = HB_ERR_INPUT;
            r.msg = "input the reference would panic on (malformed CIGAR / window descriptor / query coordinates)";
        } else if (h_terr[t] & TERR_TOO_MANY_COLS) {
            r.status = HB_ERR_CAPACITY;
            r.msg = "more than " + std::to_string(MAX_COLS_HARD) + " overlap-windows in one window";
        }
        std::vector<uint8_t> cur;
        for (uint32_t w = tg.win_begin; w < tg.win_end; w++) {
            const uint32_t len = h_outlen[w];
            if (h_nsel[w] >= 2) {
                cur.insert(cur.end(), outb + o, outb + o + len);
            } else if (!cur.empty()) {
                r.seg_len.push_back((uint32_t)cur.size());
                r.seq.insert(r.seq.end(), cur.begin(), cur.end());
                cur.clear();
            }
            o += len;
            // algorithmic bytes of the pileup build for this window (SURVEY.md §8d closed form over the
            // 31 columns the kernel consumes)
            const DevWin& dw = hbt.win[w];
            uint64_t cb = 0;
            for (uint32_t c = 0; c < h_nsel[w]; c++) {
                const DevOW& ow = hbt.ow[h_sel[(size_t)w * TOP_K + c]];
                const DevOverlap& ov = hbt.ovl[ow.ovl];
                if (ov.raw_base == RAW_NONE) cb += ow.cei - ow.csi;
                else cb += (uint64_t)ov.cig_len * W / std::max<uint32_t>(ov.tend - ov.tstart, W);  // device-windowed: its share of the CIGAR
            }
            algo += (uint64_t)(h_nsel[w] + 1) * ((dw.len + 3) / 4 + dw.len) + cb + 2ull * R_COLS * h_L[w];
        }
        if (!cur.empty()) {
            r.seg_len.push_back((uint32_t)cur.size());
            r.seq.insert(r.seq.end(), cur.begin(), cur.end());
        }
        if (r.status != HB_OK) { r.seg_len.clear(); r.seq.clear(); }
        corrected += r.seq.size();
        out_results.push_back(std::move(r));
    }
    PHASE(5);
    S.targets += nt;
    S.windows += nw;
    S.overlap_windows += hbt.ow.size();
    S.rows += total_rows;
    S.supported += n_sup;
    S.corrected_bases += corrected;
    S.kernel_launches += launches;
    S.device_launches += 1;
    S.pileup_algo_bytes += algo;

    // ---- publish: results, counters and the metadata for the debug taps / replay
    LastLaunch ll;
    ll.valid = true;
    ll.win.assign(hbt.win.begin(), hbt.win.end());
    ll.w_L.assign(h_L, h_L + nw);
    ll.w_nsel.assign(h_nsel, h_nsel + nw);
    ll.w_nsup.assign(h_nsup, h_nsup + nw);
    ll.w_rowbase.resize(nw);
    ll.w_supbase.resize(nw);
    uint64_t rb = 0, sb = 0;
    for (size_t w = 0; w < nw; w++) {
        ll.w_rowbase[w] = rb; rb += h_L[w];
        ll.w_supbase[w] = sb; sb += h_nsup[w];
        if (ctx->opt.flags & HB_FLAG_KEEP_DEBUG) ll.index[((uint64_t)hbt.win[w].rid << 32) | hbt.win[w].wid] = (uint32_t)w;
    }
    if (ctx->opt.flags & HB_FLAG_KEEP_DEBUG) {
        ll.ow_qid.resize(hbt.ow.size());
        for (size_t i = 0; i < hbt.ow.size(); i++) ll.ow_qid[i] = hbt.ovl[hbt.ow[i].ovl].qid;
    }
    ll.n_sup = n_sup;
    ll.total_rows = total_rows;
    ll.view = b;
    S.ms_worker_busy = now_ms() - t_begin;
    S.ms_worker_gpu_wait = t_wait;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        hb_stats& T = ctx->stats;
        T.ms_worker_busy += S.ms_worker_busy; T.ms_worker_gpu_wait += S.ms_worker_gpu_wait;
        T.last_launch_targets = nt; T.last_launch_windows = nw; T.last_launch_bases = corrected;
        T.targets += S.targets; T.windows += S.windows; T.overlap_windows += S.overlap_windows; T.rows += S.rows;
        T.supported += S.supported; T.corrected_bases += S.corrected_bases; T.h2d_bytes += S.h2d_bytes;
        T.d2h_bytes += S.d2h_bytes; T.kernel_launches += S.kernel_launches; T.device_launches += S.device_launches;
        T.pileup_algo_bytes += S.pileup_algo_bytes; T.gemm_flops += S.gemm_flops; T.forward_flops += S.forward_flops;
        T.ms_features += S.ms_features; T.ms_forward += S.ms_forward; T.ms_consensus += S.ms_consensus;
        for (int i = 0; i < HB_NUM_KERNEL_CLASSES; i++) { T.ms_kernel[i] += S.ms_kernel[i]; T.n_kernel[i] += S.n_kernel[i]; T.class_flops[i] += S.class_flops[i]; }
        S.ms_worker_phase[6] = now_ms() - t_mark;
        for (int i = 0; i < 8; i++) T.ms_worker_phase[i] += S.ms_worker_phase[i];
        for (auto& r : out_results) ctx->results.push_back(std::move(r));
        L->last = std::move(ll);
        ctx->last_lane = (int)(L - ctx->lanes);
    }
    return HB_OK;
}

// Hand a staged batch to the launch worker (lock held); `b` is left empty (no capacity).
// Back-pressure: at most 2 batches wait in the queue.
void enqueue_batch(hb_ctx* ctx, std::unique_lock<std::mutex>& lk, HostBatch& b, uint32_t full_targets = 0) {
    // full_targets != 0: several threads submit, and a steady-state hand-over of one of them holds up to that many targets
    if (b.tgt.empty()) return;
    // capacity hint for staging batches: the largest arrays handed over so far plus a margin, so that pinned memory is
    // allocated once per batch object and then recycled.  (No extrapolation from partial batches: a two-target
    // remainder scaled to a full launch once produced hints several times too large, and every pooled batch was then
    // re-pinned inside the next run.)
    const bool first = ctx->cap_hint[0] == 0;
    {
        size_t* h = ctx->cap_hint;
        const size_t cur[5] = {b.tgt.size(), b.win.size(), b.ovl.size(), b.ow.size(), b.cig.size()};
        // Slow start and flush remainders are smaller than a steady-state batch: scale such a batch (at least 32 targets, a fair
        // sample of the per-target sizes) up to the full hand-over size (at most 1 024 targets, at most 16x), so that the pool is
        // pinned once, at its final size, during warm-up.  (A rank of a strong-scaling job only ever hands over remainders whose
        // size depends on how its threads happened to share a step; the largest one seen so far + 25 % was exceeded inside the
        // timed region on 2 of 8 ranks: every pooled batch re-pinned, 51 cudaHostAlloc calls, 1 s.)  Never with a single
        // submitting thread: tests hand over everything in one launch.
        const double scale = (full_targets && cur[0] >= 32 && cur[0] < full_targets) ? std::min(16.0, (double)full_targets / (double)cur[0]) : 1.0;
        for (int i = 0; i < 5; i++) {
            const size_t want = (size_t)((double)cur[i] * scale);
            if (want > h[i]) h[i] = want + want / 4 + 64;
        }
    }
    if (ctx->queue.size() >= 2) {
        const auto t0 = std::chrono::steady_clock::now();
        ctx->cv_idle.wait(lk, [&] { return ctx->queue.size() < 2; });
        g_submit_wait_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
    }
    ctx->queue.push_back(std::move(b));
    b = HostBatch(ctx->device);
    ctx->cv_work.notify_one();
    if (first) {
        // The first hand-over fixes the batch geometry: create the rest of the pool now (lanes in flight + queue + one per
        // submitting thread), so that no staging batch is ever pinned in the steady state whatever the timing of the lanes.
        const uint32_t want = (uint32_t)ctx->n_lanes + 2u + std::max<uint32_t>(ctx->n_slots.load(), 4u);
        size_t h[5];
        for (int i = 0; i < 5; i++) h[i] = ctx->cap_hint[i];
        const uint32_t have = ctx->batches_alive;
        if (want > have) {
            ctx->batches_alive = want;
            lk.unlock();
            std::vector<HostBatch> fresh;
            for (uint32_t i = have; i < want; i++) {
                fresh.emplace_back(ctx->device);
                HostBatch& nb = fresh.back();
                nb.tgt.reserve(h[0]); nb.win.reserve(h[1]); nb.ovl.reserve(h[2]); nb.ow.reserve(h[3]); nb.cig.reserve(h[4]);
            }
            lk.lock();
            for (auto& nb : fresh) ctx->pool.push_back(std::move(nb));
        }
    }
}

// A staging batch with capacity: recycled from the pool, else allocated once at the largest size seen so far
// (pinned allocations are slow and serialise with the worker's CUDA calls, so growth in small steps is avoided).
void acquire_batch(hb_ctx* ctx, HostBatch& b) {
    std::unique_lock<std::mutex> lk(ctx->mu);
    size_t h[5];
    for (int i = 0; i < 5; i++) h[i] = ctx->cap_hint[i];
    if (!ctx->pool.empty()) { b = std::move(ctx->pool.back()); ctx->pool.pop_back(); }
    else { b = HostBatch(ctx->device); ctx->batches_alive++; }
    lk.unlock();
    if (h[0]) { b.tgt.reserve(h[0]); b.win.reserve(h[1]); b.ovl.reserve(h[2]); b.ow.reserve(h[3]); b.cig.reserve(h[4]); }
}

hb_ctx::ThreadSlot* my_slot(hb_ctx* ctx) {
    thread_local hb_ctx* tl_ctx = nullptr;
    thread_local hb_ctx::ThreadSlot* tl_slot = nullptr;
    thread_local uint64_t tl_gen = 0;
    if (tl_ctx == ctx && tl_slot && tl_gen == ctx->generation) return tl_slot;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const auto me = std::this_thread::get_id();
    for (auto& sl : ctx->slots)
        if (sl->owner == me) { tl_ctx = ctx; tl_slot = sl.get(); tl_gen = ctx->generation; return tl_slot; }
    ctx->slots.emplace_back(new hb_ctx::ThreadSlot{me, HostBatch(ctx->device), 0});
    ctx->n_slots.store((uint32_t)ctx->slots.size());
    tl_ctx = ctx; tl_slot = ctx->slots.back().get(); tl_gen = ctx->generation;
    return tl_slot;
}

// Record the capacities this lane ended up with; other lanes grow to them while idle (see hb_ctx::LaneSizes).  Lock held.
void publish_lane_sizes(hb_ctx* ctx, hb_ctx::Lane* L) {
    DevBuf* bufs[hb_ctx::Lane::N_DEV];
    L->all_bufs(bufs);
    bool grew = false;
    hb_ctx::LaneSizes& T = ctx->lane_sizes;
    for (int i = 0; i < hb_ctx::Lane::N_DEV; i++)
        if (bufs[i]->cap > T.dev[i]) { T.dev[i] = bufs[i]->cap; grew = true; }
    if (L->pin_small.cap > T.pin_small) { T.pin_small = L->pin_small.cap; grew = true; }
    if (L->pin_out.cap > T.pin_out) { T.pin_out = L->pin_out.cap; grew = true; }
    if (L->rows_cap > T.rows_cap) { T.rows_cap = L->rows_cap; grew = true; }
    if (grew) ctx->lane_sizes_version++;
    bool below = L->pin_small.cap < T.pin_small || L->pin_out.cap < T.pin_out || L->rows_cap < T.rows_cap;
    for (int i = 0; i < hb_ctx::Lane::N_DEV; i++) below = below || bufs[i]->cap < T.dev[i];
    if (!below) L->seen_sizes = ctx->lane_sizes_version;  // else: this lane catches up when it is next idle
}

// Grow an idle lane's buffers to the recorded sizes (no lock held; only this lane's worker touches its buffers).
void presize_lane(hb_ctx* ctx, hb_ctx::Lane* L, const hb_ctx::LaneSizes& T) {
    DevBuf* bufs[hb_ctx::Lane::N_DEV];
    L->all_bufs(bufs);
    bool ok = true;
    for (int i = 0; i < hb_ctx::Lane::N_DEV; i++) ok = ok && bufs[i]->reserve_exact(T.dev[i]) == cudaSuccess;
    ok = ok && L->pin_small.reserve_exact(T.pin_small) == cudaSuccess && L->pin_out.reserve_exact(T.pin_out) == cudaSuccess;
    if (ok && T.rows_cap > L->rows_cap) L->rows_cap = T.rows_cap;  // the six row-sized buffers are part of dev[]
    cudaStreamSynchronize(L->stream);
    if (L->last.valid) set_view_ptrs(ctx, L, L->last.view);  // buffers moved (contents copied): the kept view follows them
}


void worker_main(hb_ctx* ctx, int lane) {
    cudaSetDevice(ctx->device);
    if (ctx->have_node_cpus) pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), &ctx->node_cpus);
    hb_ctx::Lane* L = &ctx->lanes[lane];
    std::string my_err;
    t_err_sink = &my_err;
    std::unique_lock<std::mutex> lk(ctx->mu);
    for (;;) {
        ctx->cv_work.wait(lk, [&] { return ctx->stop || !ctx->queue.empty() || L->seen_sizes != ctx->lane_sizes_version; });
        if (ctx->queue.empty()) {
            if (ctx->stop) break;  // stop requested and nothing left
            // idle and another lane has grown: pre-size this lane now rather than inside its next launch
            const hb_ctx::LaneSizes T = ctx->lane_sizes;
            const uint64_t ver = ctx->lane_sizes_version;
            ctx->busy++;  // hb_flush / replay must not run while buffers move
            lk.unlock();
            presize_lane(ctx, L, T);
            lk.lock();
            L->seen_sizes = ver;
            ctx->busy--;
            ctx->cv_idle.notify_all();
            continue;
        }
        HostBatch hbt = std::move(ctx->queue.front());
        ctx->queue.pop_front();
        ctx->busy++;
        ctx->cv_idle.notify_all();
        lk.unlock();
        const int rc = run_batch(ctx, L, hbt);
        lk.lock();
        if (rc != HB_OK) {
            if (ctx->worker_rc == HB_OK) { ctx->worker_rc = rc; ctx->worker_err = my_err; }
            for (const auto& t : hbt.tgt) ctx->results.push_back(Result{t.rid, rc, {}, {}, "launch failed: " + my_err});
        } else {
            const uint64_t before = ctx->lane_sizes_version;
            publish_lane_sizes(ctx, L);
            if (ctx->lane_sizes_version != before) ctx->cv_work.notify_all();
        }
        hbt.clear();
        ctx->pool.push_back(std::move(hbt));
        ctx->busy--;
        ctx->cv_idle.notify_all();
    }
}

// Validation and the per-window bucketing of a target run on the calling (feature) thread without the
// context lock; only the final copy into the shared staging batch is serialised.
struct PreparedTarget {
    uint32_t rid, n_windows, len;
    std::vector<uint32_t> win_begin;  // [n_windows+1] CSR of the bucketed overlap-windows
    std::vector<DevOW> ow;            // bucketed by window, push order kept; ovl/win indices target-local
    uint64_t cig_bytes = 0;
    bool raw = false;                 // device windowing: `ow` is only the (overlap, window) skeleton
    std::vector<uint32_t> aln_now;    // raw: overlap-windows per alignment (0: the alignment contributes nothing, its CIGAR is not shipped)
};

// Which windows an alignment contributes to — the coordinate-only part of windowing::extract_windows
// (src/windowing.rs:53-125,260-272).  Emitted windows are the contiguous range [wa, we).
// Returns 0, or -1 where the reference would panic (inverted coordinates, a window index past the target's last window,
// the trailing-window unwrap of a None start state).
int skeleton_for_alignment(const hb_overlap& o, uint32_t W, uint32_t n_windows, uint32_t& wa, uint32_t& we) {
    wa = we = 0;
    if (o.tend < o.tstart || o.qend < o.qstart) return -1;
    if (o.tend - o.tstart < W || o.qend - o.qstart < W) return 0;          // :53-57
    const uint32_t edge = (uint32_t)(0.1f * (float)W);                      // :65
    if (o.tlen < edge) return -1;
    const uint32_t tail_thresh = o.tlen - edge;
    const uint32_t first_w = o.tstart < edge ? 0 : (o.tstart + W - 1) / W;  // :75-79
    const uint32_t last_w = o.tend > tail_thresh ? (o.tend - 1) / W + 1 : o.tend / W;  // :81-85
    if (last_w <= first_w) return 0;                                       // :106
    const bool open0 = (o.tstart % W == 0) || (o.tstart < edge);            // :120-125
    const uint32_t w_cur = o.tstart / W, w_new = o.tend / W;
    const uint32_t a = open0 ? w_cur : w_cur + 1;  // a window is emitted at every boundary crossed once a start state exists
    uint32_t e = w_new;
    if (o.tend > tail_thresh && o.tend % W != 0) {  // trailing partial window
        if (!open0 && w_new == w_cur) return -1;    // the reference unwraps a None start state here
        e = w_new + 1;
    }
    if (e <= a) return 0;
    if (e > n_windows) return -1;
    wa = a; we = e;
    return 0;
}

int prepare_target(hb_ctx* ctx, uint32_t rid, uint32_t n_windows, const hb_overlap* ovl, uint32_t n_ovl,
                   const hb_overlap_window* ow, uint32_t n_ow, PreparedTarget& P) {
    if (!ctx->have_reads) return fail(ctx, HB_ERR_STATE, "hb_upload_reads must be called before submitting targets");
    if (rid >= ctx->n_reads) return fail(ctx, HB_ERR_ARG, "rid out of range");
    const uint32_t W = ctx->opt.window_size;
    const uint32_t len = ctx->read_len[rid];
    if (n_windows != (len + W - 1) / W) return fail(ctx, HB_ERR_ARG, "n_windows != ceil(read_len / window_size)");
    if ((n_ovl && !ovl) || (n_ow && !ow)) return fail(ctx, HB_ERR_ARG, "null array");
    uint64_t cb = 0;
    for (uint32_t i = 0; i < n_ovl; i++) {
        if (ovl[i].tid != rid) return fail(ctx, HB_ERR_ARG, "overlap.tid != rid (alignments must be grouped by target)");
        if (ovl[i].qid >= ctx->n_reads) return fail(ctx, HB_ERR_ARG, "overlap.qid out of range");
        if (!ovl[i].cigar && ovl[i].cigar_len) return fail(ctx, HB_ERR_ARG, "null cigar");
        if (ovl[i].strand > 1) return fail(ctx, HB_ERR_ARG, "strand must be 0 or 1");
        cb += ovl[i].cigar_len;
    }
    P.rid = rid; P.n_windows = n_windows; P.len = len; P.cig_bytes = cb;
    P.win_begin.assign(n_windows + 1, 0);
    for (uint32_t i = 0; i < n_ow; i++) {
        if (ow[i].overlap_idx >= n_ovl || ow[i].window_idx >= n_windows) return fail(ctx, HB_ERR_ARG, "overlap_window index out of range");
        if (ow[i].cigar_end_idx < ow[i].cigar_start_idx || ow[i].cigar_end_idx > ovl[ow[i].overlap_idx].cigar_len)
            return fail(ctx, HB_ERR_ARG, "overlap_window cigar range out of bounds");
        P.win_begin[ow[i].window_idx + 1]++;
    }
    for (uint32_t w = 0; w < n_windows; w++) P.win_begin[w + 1] += P.win_begin[w];
    // bucket by window, keeping push order (= alignment order) inside each
    P.ow.resize(n_ow);
    thread_local std::vector<uint32_t> fill;  // scratch reused across calls: no allocation per target
    fill.assign(P.win_begin.begin(), P.win_begin.end() - 1);
    for (uint32_t i = 0; i < n_ow; i++) {
        const hb_overlap_window& s = ow[i];
        P.ow[fill[s.window_idx]++] = DevOW{s.overlap_idx, s.window_idx, s.tstart, s.qstart, s.qend, s.cigar_start_idx,
                                           s.cigar_start_offset, s.cigar_end_idx, s.cigar_end_offset, 0};
    }
    return HB_OK;
}

int append_target(hb_ctx* ctx, HostBatch& hbt, const PreparedTarget& P, const hb_overlap* ovl, uint32_t n_ovl) {
    const uint32_t W = ctx->opt.window_size;
    const uint32_t t_idx = (uint32_t)hbt.tgt.size();
    const uint32_t ovl_base = (uint32_t)hbt.ovl.size(), win_base = (uint32_t)hbt.win.size(), ow_base = (uint32_t)hbt.ow.size();
    const uint32_t n_ow = (uint32_t)P.ow.size();
    if (!hbt.ovl.reserve(ovl_base + n_ovl) || !hbt.cig.reserve(hbt.cig.size() + P.cig_bytes) || !hbt.win.reserve(win_base + P.n_windows) ||
        !hbt.ow.resize(ow_base + n_ow) || !hbt.tgt.reserve(t_idx + 1))
        return fail(ctx, HB_ERR_CAPACITY, "out of pinned host memory");
    for (uint32_t i = 0; i < n_ovl; i++) {
        DevOverlap d{ovl[i].qid, ovl[i].qstart, ovl[i].qend, ovl[i].strand, (uint64_t)hbt.cig.size(), ovl[i].cigar_len, t_idx,
                     ovl[i].tstart, ovl[i].tend, RAW_NONE, 0};
        if (P.raw) {
            if (P.aln_now[i] == 0) {
                d.cig_len = 0;  // contributes to no window: its CIGAR stays on the host
            } else {
                d.raw_base = (uint32_t)hbt.raw_cap;
                hbt.raw_cap += ovl[i].cigar_len / 2 + 1;
                hbt.dev_op_cap += ovl[i].cigar_len / 2 + 1 + P.aln_now[i];  // every op once + one shared op per boundary
                hbt.n_raw++;
                hbt.cig.append(ovl[i].cigar, ovl[i].cigar_len);
            }
        } else {
            hbt.cig.append(ovl[i].cigar, ovl[i].cigar_len);
        }
        hbt.ovl.push_back(d);
    }
    for (uint32_t w = 0; w < P.n_windows; w++) {
        DevWin d{};
        d.tgt = t_idx; d.rid = P.rid; d.wid = w; d.tstart = w * W;
        d.len = (w == P.n_windows - 1) ? P.len - w * W : W;
        d.ow_begin = ow_base + P.win_begin[w];
        d.ow_end = ow_base + P.win_begin[w + 1];
        hbt.win.push_back(d);
    }
    uint64_t opc = hbt.op_cap;
    for (uint32_t i = 0; i < n_ow; i++) {
        DevOW d = P.ow[i];
        d.ovl += ovl_base;
        d.win += win_base;
        if (P.raw) {
            d.op_base = 0;  // assigned on the device
        } else {
            d.op_base = (uint32_t)opc;
            opc += (d.cei - d.csi) / 2 + 1;
        }
        hbt.ow[ow_base + i] = d;
    }
    hbt.op_cap = opc;
    hbt.tgt.push_back(DevTarget{P.rid, win_base, win_base + P.n_windows, ovl_base, ovl_base + n_ovl});
    return HB_OK;
}

// Stage one prepared target in the calling thread's batch and hand the batch over when it is full.
int stage_target(hb_ctx* ctx, const PreparedTarget& P, const hb_overlap* ovl, uint32_t n_ovl) {
    hb_ctx::ThreadSlot* slot = my_slot(ctx);
    if (slot->batch.tgt.cap == 0) acquire_batch(ctx, slot->batch);
    std::string local_err;
    t_err_sink = &local_err;
    const int rc = append_target(ctx, slot->batch, P, ovl, n_ovl);
    t_err_sink = nullptr;
    if (rc) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->err = local_err; return rc; }
    // `launch_targets` is shared by the submitting threads: each stages launch_targets / n_threads targets per launch,
    // so the targets in flight (and the latency to the first launch) do not grow with the thread count
    const uint32_t lt = ctx->opt.launch_targets, ns = std::max(1u, ctx->n_slots.load(std::memory_order_relaxed));
    // ... but never less than 256 targets per launch (unless launch_targets itself is smaller): ~1 300 windows is what it takes to
    // fill 148 SMs with the one-CTA-per-window feature kernels and to amortise the ~25 launches of a batch
    uint32_t thr = std::min(lt, std::max(ctx->min_launch, lt / ns));
    // slow start: with several submitting threads, the first hand-overs after a flush are small and grow geometrically (48, 72, 108, ...
    // targets, counted over all threads), so that the GPU has work a few milliseconds after the first submit instead of after a
    // whole launch has been staged, and the threads - which fill their batches at the same rate - do not all hand over at the
    // same moment.  (A single submitting thread keeps exact launch sizes: tests and the isolated launch bench.py times rely on
    // them.)  Matters when a run is short: a rank of an 8-GPU strong-scaling job sees ~100 ms of work.
    const uint32_t full_thr = thr;
    if (ns >= 2) {
        static const uint16_t ramp[5] = {48, 72, 108, 162, 243};
        const uint32_t n = ctx->handed_total.load(std::memory_order_relaxed);
        if (n < 5) thr = std::min<uint32_t>(thr, ramp[n]);
    }
    if (slot->batch.tgt.size() >= thr) {
        std::unique_lock<std::mutex> lk(ctx->mu);
        enqueue_batch(ctx, lk, slot->batch, ns >= 2 ? std::min(full_thr, 1024u) : 0u);
        slot->handed++;
        ctx->handed_total.fetch_add(1, std::memory_order_relaxed);
    }
    return HB_OK;
}

}  // namespace

// ========================================================================================
extern "C" {

int hb_inspect_model(const char* model_path, uint32_t dims[6], uint64_t* params_hash, char* err, size_t err_cap) {
    if (!model_path || !dims) return HB_ERR_ARG;
    ModelFile mf;
    std::string e;
    const int rc = read_model_file(model_path, mf, e);
    if (rc != HB_OK) {
        if (err && err_cap) { strncpy(err, e.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
        return rc;
    }
    for (int i = 0; i < 6; i++) dims[i] = mf.cfg[3 + i];
    if (params_hash) {  // FNV-1a over the canonical tensors in name order: equal for a blob and an archive of the same weights
        uint64_t h = 1469598103934665603ull;
        for (auto& kv : mf.T) {
            for (unsigned char c : kv.first) { h ^= c; h *= 1099511628211ull; }
            const uint8_t* p = (const uint8_t*)kv.second.data();
            for (size_t i = 0; i < kv.second.size() * 4; i++) { h ^= p[i]; h *= 1099511628211ull; }
        }
        *params_hash = h;
    }
    return HB_OK;
}

const char* hb_last_error(hb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int hb_create(hb_ctx** out, int cuda_device, const char* model_path, const hb_options* opt) {
    if (!out || !model_path) { g_create_err = "null argument"; return HB_ERR_ARG; }
    *out = nullptr;
    hb_ctx* ctx = new hb_ctx();
    ctx->device = cuda_device;
    ctx->opt.struct_size = sizeof(hb_options);
    ctx->opt.window_size = 4096;
    ctx->opt.batch_size = 64;
    ctx->opt.launch_targets = 256;
    ctx->opt.flags = 0;
    if (opt) {
        if (opt->struct_size != sizeof(hb_options)) { g_create_err = "hb_options.struct_size mismatch"; delete ctx; return HB_ERR_ARG; }
        if (opt->window_size) ctx->opt.window_size = opt->window_size;
        if (opt->batch_size) ctx->opt.batch_size = opt->batch_size;
        if (opt->launch_targets) ctx->opt.launch_targets = opt->launch_targets;
        ctx->opt.flags = opt->flags;
    }
    auto bail = [&](int code) { g_create_err = ctx->err; hb_destroy(ctx); return code; };
    if (ctx->opt.window_size < 8 || ctx->opt.window_size > 8192) { ctx->err = "window_size must be in [8, 8192]"; return bail(HB_ERR_ARG); }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        ctx->err = "no CUDA device: herro_b200 has no CPU fallback";
        return bail(HB_ERR_CUDA);
    }
    if (cuda_device < 0 || cuda_device >= ndev) { ctx->err = "cuda_device out of range"; return bail(HB_ERR_ARG); }
    if (cudaSetDevice(cuda_device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(HB_ERR_CUDA); }
    if (const char* e = getenv("HERRO_B200_CHUNK_POS")) ctx->chunk_pos = (uint32_t)std::min(std::max(atoi(e), 128), 65536);
    if (const char* e = getenv("HERRO_B200_LANES")) ctx->n_lanes = std::min(std::max(atoi(e), 1), (int)hb_ctx::MAX_LANES);
    if (const char* e = getenv("HERRO_B200_MIN_LAUNCH")) ctx->min_launch = (uint32_t)std::min(std::max(atoi(e), 16), 4096);
    // debugging aids / A-B parity tests: read here once, never on the launch path
    ctx->pileup_v1 = getenv("HERRO_B200_PILEUP_V1") != nullptr;
    ctx->host_windowing = getenv("HERRO_B200_HOST_WINDOWING") != nullptr;
    if (const char* e = getenv("HERRO_B200_ARENA_ROWS")) ctx->arena_rows_per_win = (uint32_t)std::max(atoi(e), 1);
    ctx->wt.no_fuse_ln = getenv("HERRO_B200_NO_FUSE_LN") != nullptr;
    ctx->wt.no_fuse_ffn = getenv("HERRO_B200_NO_FUSE_FFN") != nullptr;
    ctx->wt.no_fuse_attn = getenv("HERRO_B200_NO_FUSE_ATTN") != nullptr;
    ctx->wt.no_fuse_oproj = getenv("HERRO_B200_NO_FUSE_OPROJ") != nullptr;
    if (!getenv("HERRO_B200_NO_NUMA_BIND")) probe_numa(ctx);
    for (int li = 0; li < ctx->n_lanes; li++) {
        auto& L = ctx->lanes[li];
        if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return bail(HB_ERR_CUDA); }
        for (auto& e : L.ev)
            if (cudaEventCreate(&e) != cudaSuccess) { ctx->err = "cudaEventCreate failed"; return bail(HB_ERR_CUDA); }
        L.bind_stream();
    }
    {   // keep freed blocks in the pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        uint64_t keep = UINT64_MAX;
        if (cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (features_configure(ctx->opt.window_size) != cudaSuccess) { ctx->err = "kernel attribute setup failed (not an sm_100a device?)"; return bail(HB_ERR_CUDA); }
    int rc = load_weights(ctx, model_path);
    if (rc) return bail(rc);
    ctx->generation = g_ctx_generation.fetch_add(1);
    for (int i = 0; i < ctx->n_lanes; i++) ctx->lanes[i].worker = std::thread(worker_main, ctx, i);
    *out = ctx;
    return HB_OK;
}

void hb_destroy(hb_ctx* ctx) {
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->stop = true;
    }
    ctx->cv_work.notify_all();
    for (auto& L : ctx->lanes) if (L.worker.joinable()) L.worker.join();
    cudaSetDevice(ctx->device);
    for (auto& L : ctx->lanes) if (L.stream) cudaStreamSynchronize(L.stream);
    for (void* p : ctx->weight_allocs) cudaFree(p);
    DevBuf* bufs[] = {&ctx->d_words, &ctx->d_word_off, &ctx->d_len, &ctx->d_qual, &ctx->d_qual_off, &ctx->d_ln};
    for (DevBuf* b : bufs) b->release();
    for (auto& L : ctx->lanes) L.release();
    ctx->pin_in.release();
    ctx->slots.clear();
    ctx->queue.clear();
    ctx->pool.clear();
    delete ctx;
}

int hb_upload_reads(hb_ctx* ctx, uint32_t n_reads, const uint64_t* const* seq_words, const uint32_t* seq_len,
                    const uint8_t* const* qual) {
    if (!ctx) return HB_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->cv_idle.wait(lk, [&] { return ctx->idle(); });
    for (auto& sl : ctx->slots)
        if (!sl->batch.tgt.empty()) return fail(ctx, HB_ERR_STATE, "hb_upload_reads with targets pending: call hb_flush first");
    if (!seq_words || !seq_len || !qual || n_reads == 0) return fail(ctx, HB_ERR_ARG, "null/empty read store");
    CK(cudaSetDevice(ctx->device));
    std::vector<uint64_t> woff(n_reads + 1, 0), qoff(n_reads + 1, 0);
    uint32_t max_len = 0;
    for (uint32_t i = 0; i < n_reads; i++) {
        woff[i + 1] = woff[i] + ((uint64_t)seq_len[i] + 31) / 32;
        qoff[i + 1] = qoff[i] + seq_len[i];
        max_len = std::max(max_len, seq_len[i]);
        if (!seq_words[i] || !qual[i]) return fail(ctx, HB_ERR_ARG, "null read");
    }
    // Padded on both sides: packed 32-base extraction may touch a few words past a read, and the pileup kernel fetches the
    // 4 bases / 4 quality bytes of a row group as whole words that may start up to 7 bytes before a read (pileup.cu).
    constexpr size_t FRONT_WORDS = 32, FRONT_QUAL = 256;  // keeps both bases 256-byte aligned
    CK(ctx->d_words.ensure((FRONT_WORDS + woff[n_reads] + 8) * 8));
    CK(cudaMemset(ctx->d_words.p, 0, FRONT_WORDS * 8));
    CK(cudaMemset(ctx->d_words.as<uint64_t>() + FRONT_WORDS + woff[n_reads], 0, 8 * 8));
    CK(ctx->d_qual.ensure(FRONT_QUAL + qoff[n_reads] + 16));
    CK(cudaMemset(ctx->d_qual.p, 33, FRONT_QUAL));
    CK(cudaMemset(ctx->d_qual.as<uint8_t>() + FRONT_QUAL + qoff[n_reads], 33, 16));
    CK(ctx->d_word_off.ensure((n_reads + 1) * 8));
    CK(ctx->d_qual_off.ensure((n_reads + 1) * 8));
    CK(ctx->d_len.ensure((size_t)n_reads * 4));
    // stage in pinned chunks
    const size_t CH = 64u << 20;
    CK(ctx->pin_in.ensure(CH));
    uint8_t* pin = ctx->pin_in.as<uint8_t>();
    auto copy_stream = [&](auto getp, auto getn, uint8_t* dbase) -> int {
        size_t fill = 0, doff = 0;
        for (uint32_t i = 0; i < n_reads; i++) {
            const uint8_t* src = (const uint8_t*)getp(i);
            size_t n = getn(i), so = 0;
            while (so < n) {
                const size_t take = std::min(n - so, CH - fill);
                memcpy(pin + fill, src + so, take);
                fill += take; so += take;
                if (fill == CH) {
                    CK(cudaMemcpy(dbase + doff, pin, fill, cudaMemcpyHostToDevice));
                    doff += fill; fill = 0;
                }
            }
        }
        if (fill) CK(cudaMemcpy(dbase + doff, pin, fill, cudaMemcpyHostToDevice));
        return HB_OK;
    };
    int rc = copy_stream([&](uint32_t i) { return (const void*)seq_words[i]; },
                         [&](uint32_t i) { return (size_t)(((uint64_t)seq_len[i] + 31) / 32 * 8); }, ctx->d_words.as<uint8_t>() + FRONT_WORDS * 8);
    if (rc) return rc;
    rc = copy_stream([&](uint32_t i) { return (const void*)qual[i]; }, [&](uint32_t i) { return (size_t)seq_len[i]; },
                     ctx->d_qual.as<uint8_t>() + FRONT_QUAL);
    if (rc) return rc;
    CK(cudaMemcpy(ctx->d_word_off.p, woff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_qual_off.p, qoff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_len.p, seq_len, (size_t)n_reads * 4, cudaMemcpyHostToDevice));
    // ln(k) table from the host libm — the value Rust's f64::ln returns (src/features.rs:507)
    ctx->ln_n = max_len + 2;
    std::vector<double> ln(ctx->ln_n);
    ln[0] = 0.0;
    for (uint32_t k = 1; k < ctx->ln_n; k++) ln[k] = std::log((double)k);
    CK(ctx->d_ln.ensure((size_t)ctx->ln_n * 8));
    CK(cudaMemcpy(ctx->d_ln.p, ln.data(), (size_t)ctx->ln_n * 8, cudaMemcpyHostToDevice));
    ctx->stats.h2d_bytes += woff[n_reads] * 8 + qoff[n_reads];
    ctx->n_reads = n_reads;
    ctx->read_len.assign(seq_len, seq_len + n_reads);
    ctx->rs = ReadStoreView{ctx->d_words.as<uint64_t>() + FRONT_WORDS, ctx->d_word_off.as<uint64_t>(), ctx->d_len.as<uint32_t>(),
                            ctx->d_qual.as<uint8_t>() + FRONT_QUAL, ctx->d_qual_off.as<uint64_t>(), n_reads};
    ctx->have_reads = true;
    return HB_OK;
}

int hb_submit_target(hb_ctx* ctx, uint32_t rid, uint32_t n_windows, const hb_overlap* ovl, uint32_t n_ovl,
                     const hb_overlap_window* ow, uint32_t n_ow) {
    if (!ctx) return HB_ERR_ARG;
    thread_local PreparedTarget P;  // its vectors keep their capacity from target to target
    P.raw = false;
    t_err_sink = nullptr;
    std::string local_err;
    {   // validation errors are written to a local string first: ctx->err is shared between feature threads
        t_err_sink = &local_err;
        const int rc = prepare_target(ctx, rid, n_windows, ovl, n_ovl, ow, n_ow, P);
        t_err_sink = nullptr;
        if (rc) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->err = local_err; return rc; }
    }
    return stage_target(ctx, P, ovl, n_ovl);
}

int hb_submit_alignments(hb_ctx* ctx, uint32_t rid, const hb_overlap* ovl, uint32_t n_ovl) {
    if (!ctx) return HB_ERR_ARG;
    auto fail_locked = [&](int code, const std::string& msg) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->err = msg; return code; };
    if (!ctx->have_reads) return fail_locked(HB_ERR_STATE, "hb_upload_reads must be called before submitting targets");
    if (rid >= ctx->n_reads) return fail_locked(HB_ERR_ARG, "rid out of range");
    if (n_ovl && !ovl) return fail_locked(HB_ERR_ARG, "null array");
    const uint32_t W = ctx->opt.window_size;
    const uint32_t n_windows = (ctx->read_len[rid] + W - 1) / W;
    if (ctx->host_windowing) {  // A-B path: extract_windows on the calling thread, then the hb_submit_target route
        std::vector<hb_overlap_window> ows;
        for (uint32_t i = 0; i < n_ovl; i++) {
            if (ovl[i].tid != rid) return fail_locked(HB_ERR_ARG, "overlap.tid != rid");
            if (host_extract_windows(ovl[i], i, W, n_windows, ows) != 0)
                return fail_locked(HB_ERR_INPUT, "malformed alignment (CIGAR / coordinates) for target " + std::to_string(rid));
        }
        return hb_submit_target(ctx, rid, n_windows, ovl, n_ovl, ows.data(), (uint32_t)ows.size());
    }
    // Device windowing (windowing_dev.cu): the host only lays out which (alignment, window) pairs exist — a function of the PAF
    // coordinates — and ships the CIGARs; a CIGAR that is malformed or disagrees with its coordinates fails this target at
    // hb_poll_corrected (HB_ERR_INPUT) instead of here.
    thread_local PreparedTarget P;  // scratch reused across calls
    thread_local std::vector<uint32_t> first, fill;
    P.raw = true;
    P.rid = rid; P.n_windows = n_windows; P.len = ctx->read_len[rid];
    P.win_begin.assign(n_windows + 1, 0);
    P.aln_now.assign(n_ovl, 0);
    first.assign(n_ovl, 0);
    uint64_t cb = 0;
    for (uint32_t i = 0; i < n_ovl; i++) {
        if (ovl[i].tid != rid) return fail_locked(HB_ERR_ARG, "overlap.tid != rid (alignments must be grouped by target)");
        if (ovl[i].qid >= ctx->n_reads) return fail_locked(HB_ERR_ARG, "overlap.qid out of range");
        if (!ovl[i].cigar && ovl[i].cigar_len) return fail_locked(HB_ERR_ARG, "null cigar");
        if (ovl[i].strand > 1) return fail_locked(HB_ERR_ARG, "strand must be 0 or 1");
        uint32_t wa, we;
        if (skeleton_for_alignment(ovl[i], W, n_windows, wa, we) != 0)
            return fail_locked(HB_ERR_INPUT, "malformed alignment (coordinates) for target " + std::to_string(rid));
        first[i] = wa;
        P.aln_now[i] = we - wa;
        if (we > wa) cb += ovl[i].cigar_len;
        for (uint32_t w = wa; w < we; w++) P.win_begin[w + 1]++;
    }
    for (uint32_t w = 0; w < n_windows; w++) P.win_begin[w + 1] += P.win_begin[w];
    P.cig_bytes = cb;
    P.ow.resize(P.win_begin[n_windows]);
    fill.assign(P.win_begin.begin(), P.win_begin.end() - 1);
    for (uint32_t i = 0; i < n_ovl; i++)  // alignment order inside every window = the reference's push order
        for (uint32_t w = first[i]; w < first[i] + P.aln_now[i]; w++) P.ow[fill[w]++] = DevOW{i, w, 0, 0, 0, 0, 0, 0, 0, 0};
    return stage_target(ctx, P, ovl, n_ovl);
}

int hb_extract_windows(const hb_overlap* ovl, uint32_t n_ovl, uint32_t window_size, uint32_t n_windows,
                       hb_overlap_window* out, uint32_t cap, uint32_t* n_out) {
    if ((!ovl && n_ovl) || !n_out || window_size == 0) return HB_ERR_ARG;
    std::vector<hb_overlap_window> v;
    for (uint32_t i = 0; i < n_ovl; i++)
        if (host_extract_windows(ovl[i], i, window_size, n_windows, v) != 0) return HB_ERR_INPUT;
    *n_out = (uint32_t)v.size();
    if (out) memcpy(out, v.data(), std::min<size_t>(v.size(), cap) * sizeof(hb_overlap_window));
    return HB_OK;
}

int hb_window_range(const hb_overlap* ovl, uint32_t window_size, uint32_t n_windows, uint32_t* first_window, uint32_t* end_window) {
    if (!ovl || !first_window || !end_window || window_size == 0) return HB_ERR_ARG;
    return skeleton_for_alignment(*ovl, window_size, n_windows, *first_window, *end_window) == 0 ? HB_OK : HB_ERR_INPUT;
}

int hb_set_launch_targets(hb_ctx* ctx, uint32_t launch_targets) {
    if (!ctx || launch_targets == 0) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->opt.launch_targets = launch_targets;
    return HB_OK;
}

int hb_set_kernel_timing(hb_ctx* ctx, int on) {
    if (!ctx) return HB_ERR_ARG;
    ctx->time_kernels.store(on != 0);
    return HB_OK;
}

int hb_flush(hb_ctx* ctx) {
    if (!ctx) return HB_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    {
        const uint32_t lt = ctx->opt.launch_targets, ns = std::max<uint32_t>(1u, (uint32_t)ctx->slots.size());
        const uint32_t full = ns >= 2 ? std::min(std::min(lt, std::max(ctx->min_launch, lt / ns)), 1024u) : 0u;
        for (auto& sl : ctx->slots) enqueue_batch(ctx, lk, sl->batch, full);  // must not race with hb_submit_* (see header)
    }
    ctx->handed_total.store(0);
    ctx->slots.clear();                                  // slots of finished feature threads are dropped;
    ctx->n_slots.store(0);
    ctx->generation = g_ctx_generation.fetch_add(1);    // live threads re-register on their next submit
    ctx->cv_idle.wait(lk, [&] { return ctx->idle(); });
    const int rc = ctx->worker_rc;
    if (rc != HB_OK) { ctx->err = ctx->worker_err; ctx->worker_rc = HB_OK; }
    return rc;
}

// Result block handed to the caller: [seg_len: n_segs u32, padded to 16][tag: u64 distance back to the block start,
// u64 magic][seq bytes].  hb_release_result finds the block from the tag in front of `seqs`, so neither call needs a
// table (or the context lock while copying).
static constexpr uint64_t RESULT_MAGIC = 0x4842524553303031ull;

int hb_poll_corrected(hb_ctx* ctx, uint32_t* rid, uint8_t** seqs, uint32_t** seg_len, uint32_t* n_segs) {
    if (!ctx || !rid || !seqs || !seg_len || !n_segs) return HB_ERR_ARG;
    Result r;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ctx->results.empty()) return 0;
        r = std::move(ctx->results.front());
        ctx->results.pop_front();
    }
    *rid = r.rid;
    *seqs = nullptr;
    *seg_len = nullptr;
    *n_segs = 0;
    if (r.status != HB_OK) {  // nothing is allocated for a failed target: there is nothing to release
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->err = "target " + std::to_string(r.rid) + ": " + r.msg;
        return r.status;
    }
    const size_t ns = r.seg_len.size();
    const size_t hdr = ((ns * 4 + 15) & ~(size_t)15) + 16;
    uint8_t* blk = (uint8_t*)malloc(hdr + r.seq.size() + 16);
    if (!blk) { std::lock_guard<std::mutex> lk(ctx->mu); return fail(ctx, HB_ERR_CAPACITY, "out of host memory"); }
    memcpy(blk, r.seg_len.data(), ns * 4);
    const uint64_t tag[2] = {(uint64_t)hdr, RESULT_MAGIC};
    memcpy(blk + hdr - 16, tag, 16);
    memcpy(blk + hdr, r.seq.data(), r.seq.size());
    *seg_len = (uint32_t*)blk;
    *seqs = blk + hdr;
    *n_segs = (uint32_t)ns;
    return 1;
}

int hb_bind_calling_thread(hb_ctx* ctx) {
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->have_node_cpus) return 0;
    return pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), &ctx->node_cpus) == 0 ? 1 : 0;
}

void hb_release_result(hb_ctx* ctx, uint8_t* seqs) {
    if (!ctx || !seqs) return;
    uint64_t tag[2];
    memcpy(tag, seqs - 16, 16);
    if (tag[1] != RESULT_MAGIC) return;  // not a block of hb_poll_corrected
    free(seqs - tag[0]);
}

int hb_get_stats(hb_ctx* ctx, hb_stats* out) {
    if (!ctx || !out) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    *out = ctx->stats;
    out->host_allocs = g_allocs.load() - ctx->alloc_base[0];
    out->ms_host_alloc = (double)(g_alloc_ns.load() - ctx->alloc_base[1]) * 1e-6;
    out->ms_submit_wait = (double)(g_submit_wait_ns.load() - ctx->alloc_base[2]) * 1e-6;
    return HB_OK;
}
int hb_reset_stats(hb_ctx* ctx) {
    if (!ctx) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->stats = hb_stats{};
    ctx->alloc_base[0] = g_allocs.load(); ctx->alloc_base[1] = g_alloc_ns.load(); ctx->alloc_base[2] = g_submit_wait_ns.load();
    return HB_OK;
}

static int find_window(hb_ctx* ctx, uint32_t rid, uint32_t wid, uint32_t* w) {
    if (!(ctx->opt.flags & HB_FLAG_KEEP_DEBUG)) return fail(ctx, HB_ERR_STATE, "context was not created with HB_FLAG_KEEP_DEBUG");
    if (ctx->last_lane < 0 || !ctx->lanes[ctx->last_lane].last.valid) return fail(ctx, HB_ERR_STATE, "no launch yet");
    const LastLaunch& last = ctx->lanes[ctx->last_lane].last;
    auto it = last.index.find(((uint64_t)rid << 32) | wid);
    if (it == last.index.end()) return fail(ctx, HB_ERR_ARG, "window not part of the most recent launch");
    *w = it->second;
    return HB_OK;
}

int hb_debug_window_shape(hb_ctx* ctx, uint32_t rid, uint32_t wid, uint32_t* shape4) {
    if (!ctx || !shape4) return HB_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->cv_idle.wait(lk, [&] { return ctx->idle(); });
    uint32_t w;
    int rc = find_window(ctx, rid, wid, &w);
    if (rc) return rc;
    const LastLaunch& last = ctx->lanes[ctx->last_lane].last;
    shape4[0] = last.w_L[w];
    shape4[1] = last.w_nsel[w];
    shape4[2] = last.w_nsup[w];
    shape4[3] = 1;
    return HB_OK;
}

int hb_debug_dump_window(hb_ctx* ctx, uint32_t rid, uint32_t wid, uint8_t* bases, uint8_t* quals, uint32_t* supported,
                         uint32_t* sup_rows, float* info_logits, float* bases_logits) {
    if (!ctx) return HB_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->cv_idle.wait(lk, [&] { return ctx->idle(); });
    CK(cudaSetDevice(ctx->device));
    uint32_t w;
    int rc = find_window(ctx, rid, wid, &w);
    if (rc) return rc;
    hb_ctx::Lane* lane = &ctx->lanes[ctx->last_lane];
    const LastLaunch& ll = lane->last;
    const uint32_t L = ll.w_L[w], ns = ll.w_nsup[w];
    const uint64_t rb = ll.w_rowbase[w], sb = ll.w_supbase[w];
    std::vector<uint8_t> tmp((size_t)L * ROW_BYTES);
    for (int pass = 0; pass < 2; pass++) {
        uint8_t* dst = pass ? quals : bases;
        if (!dst || !L) continue;
        CK(cudaMemcpy(tmp.data(), (pass ? ll.view.mat_quals : ll.view.mat_bases) + rb * ROW_BYTES, tmp.size(), cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < L; r++) memcpy(dst + (size_t)r * R_COLS, tmp.data() + (size_t)r * ROW_BYTES, R_COLS);
    }
    if (ns) {
        std::vector<uint32_t> t(ns);
        if (supported) {
            CK(cudaMemcpy(t.data(), ll.view.sup_pk + rb, (size_t)ns * 4, cudaMemcpyDeviceToHost));
            for (uint32_t k = 0; k < ns; k++) { supported[2 * k] = (t[k] >> 8) & 0xffffu; supported[2 * k + 1] = t[k] & 0xffu; }
        }
        if (sup_rows) CK(cudaMemcpy(sup_rows, ll.view.sup_row + rb, (size_t)ns * 4, cudaMemcpyDeviceToHost));
        if (info_logits) CK(cudaMemcpy(info_logits, lane->d_info.as<float>() + sb, (size_t)ns * 4, cudaMemcpyDeviceToHost));
        if (bases_logits) CK(cudaMemcpy(bases_logits, lane->d_logits.as<float>() + sb * 5, (size_t)ns * 20, cudaMemcpyDeviceToHost));
    }
    return HB_OK;
}

// ---- `herro features` dump (src/features.rs:724-764,806-839) ---------------------------------------------------------
// .npy v1.0 exactly as numpy writes it: magic, u16 header length, the dict, padded with spaces to a multiple of 64, '\n'.
static bool write_npy(const std::string& path, const std::string& descr, const std::string& shape, const void* data, size_t bytes) {
    std::string dict = "{'descr': " + descr + ", 'fortran_order': False, 'shape': " + shape + ", }";
    size_t total = 10 + dict.size() + 1;
    const size_t pad = (64 - total % 64) % 64;
    dict.append(pad, ' ');
    dict.push_back('\n');
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const uint16_t hl = (uint16_t)dict.size();
    bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&hl, 2, 1, f) == 1 && fwrite(dict.data(), 1, dict.size(), f) == dict.size() &&
              (bytes == 0 || fwrite(data, 1, bytes, f) == bytes);
    ok = fclose(f) == 0 && ok;
    return ok;
}

int hb_dump_features(hb_ctx* ctx, uint32_t rid, const char* out_dir, const char* const* read_names) {
    if (!ctx || !out_dir || !read_names) return HB_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->cv_idle.wait(lk, [&] { return ctx->idle(); });
    CK(cudaSetDevice(ctx->device));
    if (rid >= ctx->n_reads || !read_names[rid]) return fail(ctx, HB_ERR_ARG, "rid out of range / unnamed read");
    uint32_t w0;
    int rc = find_window(ctx, rid, 0, &w0);
    if (rc) return rc;
    hb_ctx::Lane* lane = &ctx->lanes[ctx->last_lane];
    const LastLaunch& ll = lane->last;
    const uint32_t W = ctx->opt.window_size, n_windows = (ctx->read_len[rid] + W - 1) / W;
    const std::string dir = std::string(out_dir) + "/" + read_names[rid];
    {   // create_dir_all
        std::string acc;
        for (size_t i = 0; i <= dir.size(); i++) {
            if (i == dir.size() || dir[i] == '/') { if (!acc.empty()) mkdir(acc.c_str(), 0777); }
            if (i < dir.size()) acc.push_back(dir[i]);
        }
    }
    static const char ASCII[13] = "ACGT*acgt#..";  // BASES_MAP inverted (src/inference.rs:23-31)
    std::vector<uint8_t> tb, tq, feat;
    std::vector<uint32_t> pk, order;
    for (uint32_t wid = 0; wid < n_windows; wid++) {
        const uint32_t w = w0 + wid;
        if (w >= ll.win.size() || ll.win[w].rid != rid || ll.win[w].wid != wid) return fail(ctx, HB_ERR_STATE, "target is not whole in the most recent launch");
        const uint32_t L = ll.w_L[w], ns = ll.w_nsup[w];
        const uint64_t rb = ll.w_rowbase[w];
        tb.resize((size_t)L * ROW_BYTES); tq.resize((size_t)L * ROW_BYTES);
        if (L) {
            CK(cudaMemcpy(tb.data(), ll.view.mat_bases + rb * ROW_BYTES, tb.size(), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(tq.data(), ll.view.mat_quals + rb * ROW_BYTES, tq.size(), cudaMemcpyDeviceToHost));
        }
        // features: [2, L', 31] u8 — plane 0 the ASCII bases, plane 1 the quality bytes
        feat.resize((size_t)2 * L * R_COLS);
        for (uint32_t r = 0; r < L; r++)
            for (int c = 0; c < R_COLS; c++) {
                feat[(size_t)r * R_COLS + c] = (uint8_t)ASCII[tb[(size_t)r * ROW_BYTES + c] < 12 ? tb[(size_t)r * ROW_BYTES + c] : 11];
                feat[(size_t)L * R_COLS + (size_t)r * R_COLS + c] = tq[(size_t)r * ROW_BYTES + c];
            }
        const std::string base = dir + "/" + std::to_string(wid);
        if (!write_npy(base + ".features.npy", "'|u1'", "(2, " + std::to_string(L) + ", " + std::to_string(R_COLS) + ")", feat.data(), feat.size()))
            return fail(ctx, HB_ERR_ARG, "cannot write " + base + ".features.npy");
        // supported: 1-D array of SupportedPos {pos: u16, ins: u8}, packed (3 bytes per record)
        pk.resize(ns);
        if (ns) CK(cudaMemcpy(pk.data(), ll.view.sup_pk + rb, (size_t)ns * 4, cudaMemcpyDeviceToHost));
        std::vector<uint8_t> rec((size_t)ns * 3);
        for (uint32_t k = 0; k < ns; k++) {
            const uint16_t pos = (uint16_t)(pk[k] >> 8);
            memcpy(&rec[(size_t)k * 3], &pos, 2);
            rec[(size_t)k * 3 + 2] = (uint8_t)(pk[k] & 0xffu);
        }
        if (!write_npy(base + ".supported.npy", "[('pos', '<u2'), ('ins', '|u1')]", "(" + std::to_string(ns) + ",)", rec.data(), rec.size()))
            return fail(ctx, HB_ERR_ARG, "cannot write " + base + ".supported.npy");
        // ids: the query reads of ALL surviving overlaps of the window in final rank order (src/features.rs:569)
        uint32_t n1 = 0;
        CK(cudaMemcpy(&n1, ll.view.w_n1 + w, 4, cudaMemcpyDeviceToHost));
        order.resize(n1);
        if (n1) CK(cudaMemcpy(order.data(), ll.view.rank_ow + ll.win[w].ow_begin, (size_t)n1 * 4, cudaMemcpyDeviceToHost));
        FILE* f = fopen((base + ".ids.txt").c_str(), "wb");
        if (!f) return fail(ctx, HB_ERR_ARG, "cannot write " + base + ".ids.txt");
        for (uint32_t k = 0; k < n1; k++) {
            const uint32_t q = order[k] < ll.ow_qid.size() ? ll.ow_qid[order[k]] : 0;
            fprintf(f, "%s\n", q < ctx->n_reads && read_names[q] ? read_names[q] : "?");
        }
        fclose(f);
    }
    return HB_OK;
}

int hb_selftest_gemm(int cuda_device, uint32_t M, uint32_t N, uint32_t K, int act, int res, uint32_t lda_extra,
                     float* max_abs_err, float* max_abs_ref, float* ms_tc, float* ms_simt) {
    if (!max_abs_err || !max_abs_ref || M % 128 || N % 128 || K % 64 || (res && act && act != 3)) return HB_ERR_ARG;
    if (cudaSetDevice(cuda_device) != cudaSuccess) return HB_ERR_CUDA;
    const size_t lda = (size_t)K + lda_extra;
    std::vector<float> hA((size_t)M * lda), hW((size_t)N * K), hb(N), hR((size_t)M * N);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0); };
    for (auto& v : hA) v = rnd();
    for (auto& v : hW) v = rnd() * 0.1f;
    for (auto& v : hb) v = rnd();
    for (auto& v : hR) v = rnd();
    float *dA, *dW, *db, *dR, *dC1, *dC2;
    void *hi = nullptr, *lo = nullptr;
    bool ok = cudaMalloc(&dA, hA.size() * 4) == cudaSuccess && cudaMalloc(&dW, hW.size() * 4) == cudaSuccess &&
              cudaMalloc(&db, hb.size() * 4) == cudaSuccess && cudaMalloc(&dR, hR.size() * 4) == cudaSuccess &&
              cudaMalloc(&dC1, hR.size() * 4) == cudaSuccess && cudaMalloc(&dC2, hR.size() * 4) == cudaSuccess;
    if (!ok) return HB_ERR_CUDA;
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dR, hR.data(), hR.size() * 4, cudaMemcpyHostToDevice);
    if (split_weights(dW, hW.size(), &hi, &lo) != cudaSuccess) return HB_ERR_CUDA;
    void *ahi = nullptr, *alo = nullptr, *ohi = nullptr, *olo = nullptr;
    if (split_weights(dA, hA.size(), &ahi, &alo) != cudaSuccess) return HB_ERR_CUDA;
    if (cudaMalloc(&ohi, hR.size() * 2) != cudaSuccess || cudaMalloc(&olo, hR.size() * 2) != cudaSuccess) return HB_ERR_CUDA;
    int num_sms = 148;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, cuda_device);
    GemmArgs ga{};
    ga.Ahi = (const __nv_bfloat16*)ahi; ga.Alo = (const __nv_bfloat16*)alo; ga.lda = lda;
    ga.Whi = (const __nv_bfloat16*)hi; ga.Wlo = (const __nv_bfloat16*)lo; ga.K = K;
    ga.bias = db; ga.out = dC2; ga.res = res ? dR : nullptr; ga.ldc = N;
    ga.out_hi = (__nv_bfloat16*)ohi; ga.out_lo = (__nv_bfloat16*)olo; ga.ldo = N;
    ga.m_tiles = M / 128; ga.n_chunks = N / 128; ga.k_blocks = K / 64;
    ga.mode = res ? GEMM_OUT_F32_RES : (act == 2 ? GEMM_OUT_SPLIT_RELU : (act ? GEMM_OUT_F32_RELU : GEMM_OUT_F32));
    if (act == 3) {  // residual + fused LayerNorm epilogue (N must be 128); compares the fp32 residual-stream output
        if (N != 128) return HB_ERR_ARG;
        ga.mode = GEMM_OUT_F32_RES_LN; ga.res = dC2; ga.ln_g = db; ga.ln_b = db; res = 1; act = 0;
        cudaMemcpy(dC2, dR, hR.size() * 4, cudaMemcpyDeviceToDevice);
    }
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2; rep++) {  // second repetition is the timed one
        if (ga.mode == GEMM_OUT_F32_RES_LN) cudaMemcpy(dC2, dR, hR.size() * 4, cudaMemcpyDeviceToDevice);  // in-place residual
        cudaEventRecord(e0);
        gemm_simt(act ? 1 : 0, res, dA, (int)lda, dW, db, dC1, (int)N, res ? dR : nullptr, M, (int)N, (int)K, 0);
        cudaEventRecord(e1);
        e = gemm_tc(ga, num_sms, 0);
        cudaEventRecord(e2);
        if (e != cudaSuccess) break;
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) break;
    }
    if (e == cudaSuccess && act == 2) {  // recombine the split output into dC2 for the comparison
        std::vector<uint16_t> h1(hR.size()), h2(hR.size());
        cudaMemcpy(h1.data(), ohi, h1.size() * 2, cudaMemcpyDeviceToHost);
        cudaMemcpy(h2.data(), olo, h2.size() * 2, cudaMemcpyDeviceToHost);
        std::vector<float> c(hR.size());
        for (size_t i = 0; i < c.size(); i++) {
            uint32_t a = (uint32_t)h1[i] << 16, b2 = (uint32_t)h2[i] << 16;
            float fa, fb;
            memcpy(&fa, &a, 4); memcpy(&fb, &b2, 4);
            c[i] = fa + fb;
        }
        cudaMemcpy(dC2, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
    }
    int rc = HB_OK;
    if (e != cudaSuccess) {
        g_create_err = std::string("selftest: ") + cudaGetErrorString(e);
        rc = HB_ERR_CUDA;
    } else {
        std::vector<float> c1(hR.size()), c2(hR.size());
        cudaMemcpy(c1.data(), dC1, c1.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(c2.data(), dC2, c2.size() * 4, cudaMemcpyDeviceToHost);
        float me = 0, mr = 0;
        for (size_t i = 0; i < c1.size(); i++) {
            float d = fabsf(c1[i] - c2[i]);
            if (!(d <= me)) me = d;  // NaN propagates
            mr = std::max(mr, fabsf(c1[i]));
        }
        *max_abs_err = me;
        *max_abs_ref = mr;
        if (ms_simt) cudaEventElapsedTime(ms_simt, e0, e1);
        if (ms_tc) cudaEventElapsedTime(ms_tc, e1, e2);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dR); cudaFree(dC1); cudaFree(dC2); cudaFree(hi); cudaFree(lo); cudaFree(ahi); cudaFree(alo); cudaFree(ohi); cudaFree(olo);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return rc;
}

int hb_replay_last_launch(hb_ctx* ctx, uint32_t iters, float* ms) {
    if (!ctx || !ms) return HB_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->cv_idle.wait(lk, [&] { return ctx->idle(); });
    CK(cudaSetDevice(ctx->device));
    if (ctx->last_lane < 0 || !ctx->lanes[ctx->last_lane].last.valid) return fail(ctx, HB_ERR_STATE, "no launch to replay");
    hb_ctx::Lane* L = &ctx->lanes[ctx->last_lane];
    const BatchView b = L->last.view;
    uint64_t launches = 0;
    L->kt.on = false;
    L->kt.st = L->stream;
    CK(cudaStreamSynchronize(L->stream));
    CK(cudaEventRecord(L->ev[6], L->stream));
    for (uint32_t it = 0; it < iters; it++) {
        int rc = zero_scratch(ctx, L, b);
        if (rc) return rc;
        launches += launch_features_a(b, L->stream, L->kt);
        launches += launch_pileup(b, L->stream, L->kt, ctx->pileup_v1);
        launches += launch_features_c1(b, L->stream, L->kt);
        rc = launch_tail(ctx, L, b, L->last.n_sup, &launches);
        if (rc) return rc;
    }
    CK(cudaEventRecord(L->ev[7], L->stream));
    CK(cudaStreamSynchronize(L->stream));
    CK(cudaEventElapsedTime(ms, L->ev[6], L->ev[7]));
    L->kt.discard();
    ctx->stats.kernel_launches += launches;
    return HB_OK;
}

}  // extern "C"
