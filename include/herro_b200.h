/* herro_b200.h — C ABI of the B200-native HERRO hot path (features -> inference -> consensus).
 *
 * The reference (lbcb-sci/herro @ 9cd0296) has no FFI layer: the three stages are Rust
 * worker threads joined by crossbeam channels.  This header is what a `mod ffi` in the Rust
 * host binds so that `herro inference` keeps its CLI, FASTQ loading, `--read-alns` batches,
 * windowing and FASTA writer (src/main.rs, src/lib.rs, src/haec_io.rs, src/overlaps.rs,
 * src/windowing.rs) and replaces
 *
 *   features::extract_features      src/features.rs:326-583   (called at src/lib.rs:176-183)
 *   inference::inference_worker     src/inference.rs:177-212  (spawned at src/lib.rs:189-196)
 *   consensus::consensus_worker     src/consensus.rs:229-263  (spawned at src/lib.rs:198-199)
 *
 * with calls into this library (INTEGRATION.md shows the Rust side).  Plain pointers and
 * sizes only; no torch types; never unwinds or aborts (the reference is `panic = "abort"`,
 * Cargo.toml:14-16): every entry point returns HB_OK or a negative hb_status, and
 * hb_last_error() gives the message.
 *
 * Threading: hb_submit_* may be called concurrently from the host's feature threads
 * (`-t`, src/lib.rs:159-187); hb_poll_corrected / hb_release_result from one thread per
 * context (the former consensus thread feeding correction_writer, src/lib.rs:267-291).
 * One context per GPU (`-d`), like the reference's per-device worker group.
 */
#ifndef HERRO_B200_H
#define HERRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 1

typedef enum hb_status {
    HB_OK = 0,
    HB_ERR_ARG = -1,      /* bad argument / NULL pointer / out-of-range id                      */
    HB_ERR_CUDA = -2,     /* CUDA runtime failure (no device, OOM, launch error)                */
    HB_ERR_MODEL = -3,    /* weights file missing / malformed / unsupported dimensions          */
    HB_ERR_INPUT = -4,    /* input on which the reference itself would panic (bad CIGAR, ...)   */
    HB_ERR_CAPACITY = -5, /* an internal capacity was exceeded and could not be grown (e.g. more   */
                          /* than 60000 overlaps covering one window, out of memory)            */
    HB_ERR_STATE = -6     /* call sequence error (e.g. submit before hb_upload_reads)           */
} hb_status;

/* Overlap + CIGAR of one alignment — `struct Overlap` / `struct Alignment`
 * (src/overlaps.rs:44-55,91-95).  strand: 0 = '+', 1 = '-'.  cigar: ASCII `[0-9]+[MID]`
 * exactly as parsed from the PAF `cg:Z:` field (src/overlaps.rs:172); caller-owned, copied
 * before the call returns. */
typedef struct hb_overlap {
    uint32_t qid, qlen, qstart, qend;
    uint32_t strand;
    uint32_t tid, tlen, tstart, tend;
    const uint8_t* cigar;
    uint32_t cigar_len;
} hb_overlap;

/* One OverlapWindow (src/windowing.rs:6-16) plus the window it was pushed to
 * (`windows[i].push(...)`, src/windowing.rs:161,233,262).  cigar_*_idx are BYTE offsets into
 * the overlap's CIGAR, cigar_*_offset are base counts from the start of that op
 * (SURVEY.md App. G). */
typedef struct hb_overlap_window {
    uint32_t overlap_idx; /* index into the `ovl` array of the same hb_submit_target call */
    uint32_t window_idx;  /* 0 .. n_windows-1                                             */
    uint32_t tstart, qstart, qend;
    uint32_t cigar_start_idx, cigar_start_offset, cigar_end_idx, cigar_end_offset;
} hb_overlap_window;

typedef struct hb_options {
    uint32_t struct_size;    /* = sizeof(hb_options)                                                     */
    uint32_t window_size;    /* `-w` (src/main.rs:69-74), default 4096                                    */
    uint32_t batch_size;     /* `-b` (src/main.rs:98-102): the reference groups consecutive windows of a  */
                             /* read in chunks of b and pads each chunk to its longest window             */
                             /* (src/features.rs:884-893, src/inference.rs:73-97); that padding is part   */
                             /* of the model input, so it is reproduced per window       */
    uint32_t launch_targets; /* target reads per device launch (cross-read batching); 0 = default         */
    uint32_t flags;          /* HB_FLAG_*                                                                */
} hb_options;

#define HB_FLAG_KEEP_DEBUG 1u /* keep per-window intermediates of the last launch for hb_debug_*        */

typedef struct hb_ctx hb_ctx; /* one per GPU: streams, weights, read-store replica, staging buffers */

/* Create a context on `cuda_device` and load the forward weights (replaces tch::CModule::load_on_device,
 * src/inference.rs:185).  `model_path` is what `-m` names: a TorchScript archive (`torch.jit.save`; ZIP with stored
 * entries + data.pkl, read natively, no libtorch) of a module with the architecture this library implements
 * (oracle/forward_ref.py naming; an optional `stem_bn` is folded), or the HB200W1 blob of herro_b200/weights.py.
 * Any other graph is HB_ERR_MODEL with the first missing parameter named: TorchScript code is not executed. */
int hb_create(hb_ctx** out, int cuda_device, const char* model_path, const hb_options* opt);
void hb_destroy(hb_ctx* ctx);

/* Architecture of a model file without creating a context (host only, no CUDA call): dims = stem_k, channels, heads,
 * layers, ffn, collapse; *params_hash (may be NULL) = FNV-1a-64 over the canonical fp32 tensors in name order, equal for
 * a blob and an archive holding the same weights.  `err` (may be NULL) receives the message on failure. */
int hb_inspect_model(const char* model_path, uint32_t dims[6], uint64_t* params_hash, char* err, size_t err_cap);

/* Replicate the read store on the GPU.  Layout is HAECRecord verbatim (src/haec_io.rs:19-24,
 * 77-81): seq_words[i] = 2-bit little-endian packing, 32 bases per u64, A0 C1 G2 T3
 * (src/haec_io.rs:121-136); seq_len[i] bases; qual[i] = raw Phred+33 bytes (seq_len[i] of them).
 * Read ids are positions in this array (`rid`), as in the reference (src/overlaps.rs:332-336). */
int hb_upload_reads(hb_ctx* ctx, uint32_t n_reads, const uint64_t* const* seq_words, const uint32_t* seq_len,
                    const uint8_t* const* qual);

/* Submit one target read: the replacement of extract_features(rid, reads, overlaps, ...)
 * (src/features.rs:326-333) for a host that still runs windowing::extract_windows itself.
 * n_windows = ceil(len/W) (src/features.rs:338).  `ow` lists the OverlapWindows in the order
 * extract_windows pushed them (alignment order), each tagged with its window.  Every
 * overlap must have tid == rid (src/overlaps.rs:189-192).  Results become available through
 * hb_poll_corrected after enough targets were submitted or after hb_flush. */
int hb_submit_target(hb_ctx* ctx, uint32_t rid, uint32_t n_windows, const hb_overlap* ovl, uint32_t n_ovl,
                     const hb_overlap_window* ow, uint32_t n_ow);

/* Same, but the library also performs windowing::extract_windows (src/windowing.rs:44-273)
 * on the raw alignments — for hosts that hand over `(tid, Vec<Alignment>)` straight from
 * alignment_reader (src/overlaps.rs:371-373).  The windowing runs ON THE DEVICE: the calling
 * thread only derives, from the PAF coordinates, which windows each alignment contributes to, and copies the CIGARs;
 * no CIGAR byte is parsed on the host.  Consequently a malformed CIGAR (or one that does not span its PAF
 * coordinates) is reported for its target by hb_poll_corrected (HB_ERR_INPUT), not by this call; coordinate errors
 * (inverted ranges, windows past the end of the target) still fail here. */
int hb_submit_alignments(hb_ctx* ctx, uint32_t rid, const hb_overlap* ovl, uint32_t n_ovl);

/* Host-only utility: the windows [*first_window, *end_window) one alignment contributes OverlapWindows to — the
 * coordinate-only part of extract_windows (src/windowing.rs:53-125,260-272).  HB_ERR_INPUT where the reference panics. */
int hb_window_range(const hb_overlap* ovl, uint32_t window_size, uint32_t n_windows, uint32_t* first_window, uint32_t* end_window);

/* Host-only utility (no context, no GPU): windowing::extract_windows (src/windowing.rs:44-273) of
 * the n_ovl alignments of one target that has n_windows windows (overlap_idx = position in
 * `ovl`).  Writes up to `cap` records in push order, returns the number produced in *n_out
 * (may exceed cap: call again with a larger buffer); HB_ERR_INPUT if the reference would panic. */
int hb_extract_windows(const hb_overlap* ovl, uint32_t n_ovl, uint32_t window_size, uint32_t n_windows,
                       hb_overlap_window* out, uint32_t cap, uint32_t* n_out);

/* Change hb_options.launch_targets (targets per device launch, staged per feature thread) for subsequent
 * submissions.  Call between hb_flush and the next hb_submit_*. */
int hb_set_launch_targets(hb_ctx* ctx, uint32_t launch_targets);

/* Per-kernel CUDA-event timing (hb_stats.ms_kernel / n_kernel) for subsequent launches: off by default, because
 * the two event records around each of the ~25 kernels of a launch cost device time; bench.py switches it on
 * for the one isolated launch its roofline is computed from.  Stage-level times (ms_features, ...) are always kept. */
int hb_set_kernel_timing(hb_ctx* ctx, int on);

/* Launch whatever is pending (every feature thread stages its own batch) and wait until every submitted
 * target has a result queued.  Call it once the submitting threads are quiescent (the reference's
 * equivalent is the alignment channel closing, src/lib.rs:186); it must not race with hb_submit_*. */
int hb_flush(hb_ctx* ctx);

/* Pop one finished target: the `(rid, Vec<Vec<u8>>)` of src/consensus.rs:253-257.
 * Returns 1 and fills the outputs if a result was popped, 0 if none is queued, <0 if the popped target
 * failed: then *rid names it, *seqs / *seg_len are NULL, *n_segs is 0 (nothing to release), hb_last_error()
 * says why (HB_ERR_INPUT: the reference would have panicked on its alignments; HB_ERR_CAPACITY / HB_ERR_CUDA:
 * its launch failed), and the other targets are unaffected — a host that wants the reference's behaviour
 * aborts, one that wants to keep going logs the read and polls on.
 * `*seqs` = the segments back to back (ASCII ACGT), `*seg_len[k]` their lengths; n_segs == 0
 * means consensus() returned None or produced nothing — the read is omitted from the FASTA
 * (src/consensus.rs:95-98, src/lib.rs:282-288).  Buffers stay valid until
 * hb_release_result(ctx, *seqs). */
int hb_poll_corrected(hb_ctx* ctx, uint32_t* rid, uint8_t** seqs, uint32_t** seg_len, uint32_t* n_segs);
void hb_release_result(hb_ctx* ctx, uint8_t* seqs);

const char* hb_last_error(hb_ctx* ctx); /* ctx may be NULL: error of a failed hb_create */

/* Bind the calling thread to the CPUs of the NUMA node the context's GPU is attached to (the library's own
 * launch workers always are).  For the host's feature / consumer threads of this device (src/lib.rs:159-199) on
 * multi-socket boxes.  Returns 1 if bound, 0 if the topology is unknown (nothing changed), <0 on error. */
int hb_bind_calling_thread(hb_ctx* ctx);

/* ---- counters (the reference only has progress bars, src/pbars.rs) -------------------- */
#define HB_NUM_KERNEL_CLASSES 16
/* kernel classes of ms_kernel[] / n_kernel[] */
enum { HB_K_TOKENIZE = 0, HB_K_PASS1, HB_K_SCORES, HB_K_PASS2A, HB_K_SCAN, HB_K_PILEUP, HB_K_LISTS, HB_K_STEM,
       HB_K_LAYERNORM, HB_K_GEMM, HB_K_ATTENTION, HB_K_HEADS, HB_K_CONSENSUS,
       HB_K_FFN /* fused FFN kernel */, HB_K_QKV_ATTN /* fused QKV projection + attention kernel */ };
typedef struct hb_stats {
    uint64_t targets, windows, overlap_windows, rows, supported, corrected_bases;
    uint64_t h2d_bytes, d2h_bytes, kernel_launches, device_launches /* batches */;
    uint64_t pileup_algo_bytes; /* algorithmic bytes of the pileup-build kernel */
    uint64_t gemm_flops;        /* algorithmic FLOPs of the dense contractions (QKV/out/FFN/collapse)     */
    uint64_t forward_flops;     /* all forward FLOPs at the supported positions                            */
    double ms_features, ms_forward, ms_consensus;  /* CUDA-event time per stage, summed over launches      */
    double ms_kernel[HB_NUM_KERNEL_CLASSES];       /* CUDA-event time per kernel class (timed launches)    */
    uint64_t n_kernel[HB_NUM_KERNEL_CLASSES];      /* launches per kernel class                            */
    uint64_t last_launch_targets, last_launch_windows, last_launch_bases; /* what hb_replay_last_launch re-runs */
    double ms_worker_busy;   /* host wall time the launch worker spent inside launches (staging+GPU+copy-back) */
    double ms_worker_gpu_wait; /* of which: blocked in stream synchronisation                               */
    uint64_t host_allocs;    /* page-locked / device allocations made since the last reset (0 in steady state)  */
    double ms_host_alloc;    /* wall time spent in them                                                         */
    double ms_submit_wait;   /* wall time hb_submit_* callers were blocked on back-pressure (summed over threads) */
    uint64_t class_flops[HB_NUM_KERNEL_CLASSES]; /* algorithmic FLOPs (31 read tokens per supported position) by kernel class */
    double ms_worker_phase[8]; /* launch-worker wall time by phase: 0 buffers+H2D enqueue, 1 feature launches,
                                  2 wait (row counts), 3 forward+consensus launches, 4 wait (results),
                                  5 per-read reassembly, 6 publish                                               */
} hb_stats;
int hb_get_stats(hb_ctx* ctx, hb_stats* out);
int hb_reset_stats(hb_ctx* ctx);

/* ---- parity taps (need HB_FLAG_KEEP_DEBUG; valid for targets of the most recent launch) -- */
/* shape4 = L' (rows), n_alns, n_supported, has_logits */
int hb_debug_window_shape(hb_ctx* ctx, uint32_t rid, uint32_t wid, uint32_t* shape4);
/* bases/quals: [L',31] u8 (tokens of BASES_MAP src/inference.rs:23-31 / raw quals), i.e. the
 * arguments of FeaturesOutput::update after prepare_examples; supported: [n,2] u32 (pos, ins);
 * sup_rows: [n] u32; info_logits [n], bases_logits [n,5] f32.  Any pointer may be NULL. */
int hb_debug_dump_window(hb_ctx* ctx, uint32_t rid, uint32_t wid, uint8_t* bases, uint8_t* quals,
                         uint32_t* supported, uint32_t* sup_rows, float* info_logits, float* bases_logits);

/* `herro features` (src/lib.rs:50-111, src/features.rs:724-764,806-839) from the device path: for target `rid` of the
 * most recent launch (HB_FLAG_KEEP_DEBUG) writes, per window, under <out_dir>/<read_names[rid]>/ :
 *   <wid>.features.npy   u8 [2, L', 31]: plane 0 the pileup as raw ASCII (ACGT / acgt / '*' '#' / '.'), plane 1 the qualities
 *   <wid>.supported.npy  1-D records {pos: u16, ins: u8}  (SupportedPos, src/features.rs:894-898)
 *   <wid>.ids.txt        the query read ids of all overlaps that survived the filter, in final rank order (src/features.rs:569)
 * i.e. the arguments of FeaturesOutput::update.  read_names[i] = id of read i (NUL-terminated), n_reads entries.
 * The .npy headers are the ones numpy writes (v1.0, padded to 64 bytes). */
int hb_dump_features(hb_ctx* ctx, uint32_t rid, const char* out_dir, const char* const* read_names);

/* ---- device-resident replay, used by bench.py for the HBM-resident `value` -------------- */
/* Re-run all device stages of the most recent launch from its inputs already in HBM
 * (no host<->device copies), `iters` times; returns the CUDA-event milliseconds in *ms. */
int hb_replay_last_launch(hb_ctx* ctx, uint32_t iters, float* ms);

/* ---- diagnostics ------------------------------------------------------------------------- */
/* Runs the tcgen05 (bf16x3) contraction kernel and the fp32 SIMT one on the same random
 * [M,K]x[N,K]^T problem (M % 128 == 0, N % 128 == 0, K % 64 == 0; lda = K + lda_extra; act: 0 none,
 * 1 ReLU, 2 ReLU with split-bf16 output; res: add a residual) and reports the largest absolute
 * difference, the largest |reference| and both kernel times. */
int hb_selftest_gemm(int cuda_device, uint32_t M, uint32_t N, uint32_t K, int act, int res, uint32_t lda_extra,
                     float* max_abs_err, float* max_abs_ref, float* ms_tc, float* ms_simt);

#ifdef __cplusplus
}
#endif
#endif /* HERRO_B200_H */
