/*
 * specdec_b200.h -- C ABI of libspecdec_b200.so (sm_100a), the drop-in boundary for the
 * speculative-sampling verify hot path of dadiaokua/speculative-decoding.
 *
 * The reference has no FFI (it is pure Python over torch); each entry point
 * below names the reference Python interface it replaces (paths relative to the reference
 * tree).  INTEGRATION.md shows the ctypes / torch.library binding a maintainer adds.
 *
 * Conventions: every function returns 0 on success, <0 for an invalid argument
 * (SPECDEC_ERR_*), >0 for a cudaError_t.  Nothing throws, nothing synchronises: all work is enqueued on `stream`.
 * Device memory is never allocated (the caller passes workspace; the n-gram tables are allocated by their create
 * call).  ONE lazy host-side allocation exists: the first specdec_verify call of a device that takes the two-chunk
 * pipeline (16-bit logits, B >= 128, plain modes) creates the library's auxiliary non-blocking stream and its
 * fork/join events (cudaStreamCreateWithPriority + 9 cudaEventCreate, ~50 us, once per device, kept until exit).
 * Thread safety: the entry points that enqueue work or change options serialise on one process-wide mutex for the
 * duration of the enqueue; options (specdec_set_option) and profiling events are process-wide, not per call.  All pointers
 * are DEVICE pointers unless named host_*.  Strides are in ELEMENTS.  Logits are read-only
 * (unlike TopKProcessor._process, utils/logits_processor.py:62, which mutates its input).
 */
#ifndef SPECDEC_B200_H
#define SPECDEC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* specdec_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define SPECDEC_API __attribute__((visibility("default")))
#else
#define SPECDEC_API
#endif

/* logits dtypes */
#define SPECDEC_F32 0
#define SPECDEC_BF16 1
#define SPECDEC_F16 2

/* sample_mode: LogitsProcessor.sample() semantics */
#define SPECDEC_SAMPLE_GREEDY 0 /* GreedyProcessor.sample: argmax, first index (utils/logits_processor.py:35-36) */
#define SPECDEC_SAMPLE_INVCDF 1 /* Multinomial/TopK/Nucleus sample() restated as inverse CDF on an injected uniform */

/* flags */
#define SPECDEC_ACCEPT_BATCHED 1  /* accept iff u < (q<=0 ? 1 : min(1,p/q))  engine/infer_engine.py:303-305;
                                     default: reject iff u > p/q             sampling/speculative_decoding.py:139-145 */
#define SPECDEC_NO_BONUS 2        /* no bonus row; next_token=-1 when all accepted (engine/infer_engine.py:276) */
#define SPECDEC_SKIP_ADJUST 4     /* skip_sample_adjustment (sampling/speculative_decoding.py:167-170) */
#define SPECDEC_NGRAM 8           /* accept iff draft==sample(p_i); no drafter logits (ngram_assisted/ngram_assisted.py:114-141) */
#define SPECDEC_RESID_FALLBACK 16 /* residual mass <= 1e-12 -> sample from p (engine/infer_engine.py:319-321) */
#define SPECDEC_OFFSET_DEVICE 32  /* philox_offset is the ADDRESS of a uint64 in device memory; the kernels read the word
                                     when they run, so a captured CUDA graph / device-resident decode loop
                                     (engine/infer_engine.py:211-338 without host scalars) advances its uniforms by
                                     bumping that word.  specdec_sample_rows: same meaning, signalled by
                                     lane_id | SPECDEC_LANE_OFFSET_DEVICE */
#define SPECDEC_LANE_OFFSET_DEVICE (1 << 30)

#define SPECDEC_ERR_ARG (-1)
#define SPECDEC_ERR_WORKSPACE (-2)
#define SPECDEC_ERR_DTYPE (-3)
#define SPECDEC_ERR_RANGE (-4)

SPECDEC_API int specdec_version(void);
SPECDEC_API const char* specdec_error_string(int code);

/* Workspace sizes (bytes) for the three row-processing entry points. */
SPECDEC_API size_t specdec_workspace_bytes(int64_t rows);          /* specdec_process_probs */
SPECDEC_API size_t specdec_verify_workspace_bytes(int B, int gamma, int V);
SPECDEC_API size_t specdec_sample_rows_workspace_bytes(int64_t rows, int V);

/*
 * One speculative verify step for B sequences: replaces, per sequence,
 *   p = logits_processor(target_logits[:, cp-1:cp+gamma-1])   sampling/speculative_decoding.py:135-136
 *   q[0,k] = logits_processor(draft_logits_k)                  sampling/speculative_decoding.py:107,120-122
 *   r = rand(gamma); fractions = p/q; first rejection n         :139-145
 *   stop-token scan                                             :150-155
 *   bonus (n==gamma) / max_fn(p[n]-q[n]) / skip adjustment; x = sample(p_p)   :158-171
 * and, with SPECDEC_ACCEPT_BATCHED|SPECDEC_NO_BONUS|SPECDEC_RESID_FALLBACK, the batched
 * loop body engine/infer_engine.py:276-336.  The processor is (temperature, top_k, top_p):
 * top_k<=0 or >=V disables top-k, top_p<=0 or >=1 disables nucleus
 * (utils/logits_processor.py:13-15,59-63,73-81,92-103).
 *
 * target_logits [B, gamma+1, V] (row gamma = bonus row; [B,gamma,V] suffices with NO_BONUS),
 * draft_logits  [B, gamma, V]  (NULL with SPECDEC_NGRAM), draft_tokens [B,gamma] int64.
 * u_accept [B,gamma], u_sample [B] in [0,1): if NULL they are Philox4x32-10 uniforms keyed
 * by (philox_seed, philox_offset, seq_id0+b, position) -- independent of sharding.
 * Outputs: n_accepted[B], next_token[B] (-1 if none), accept_mask[B,gamma] (per-position test,
 * also past the first rejection), p_tok/q_tok[B,gamma] = P_i[tok_i], Q_i[tok_i],
 * first_stop[B] = first accepted draft that is a stop token or -1,
 * next_prob[B] (nullable) = probability of next_token under the distribution it was drawn from
 * when that distribution is a processed target row, else 0,
 * packed[B, gamma+2] int32 (nullable) = {n, tok_0..tok_{n-1}, next_token, -1...} for the
 * multi-GPU all-gather.
 */
SPECDEC_API int specdec_verify(const void* target_logits, const void* draft_logits, int dtype,
                   const int64_t* draft_tokens, const float* u_accept, const float* u_sample,
                   uint64_t philox_seed, uint64_t philox_offset, int64_t seq_id0,
                   int B, int gamma, int V,
                   int64_t stride_tb, int64_t stride_tg, int64_t stride_db, int64_t stride_dg,
                   float temperature, int top_k, float top_p, int sample_mode, int flags,
                   const int64_t* stop_tokens, int n_stop,
                   int32_t* n_accepted, int64_t* next_token, uint8_t* accept_mask,
                   float* p_tok, float* q_tok, int32_t* first_stop, float* next_prob, int32_t* packed,
                   void* workspace, size_t workspace_bytes, specdec_stream_t stream);

/* Measurement hook (bench.py): when non-NULL, specdec_verify records these cudaEvent_t on its stream
 * before the row-statistics kernel, between the two kernels and after the decide kernel. */
SPECDEC_API int specdec_set_profile_events(void* ev_start, void* ev_mid, void* ev_end);
/* Test / tuning hooks (defaults in brackets):
 *  "force_ldg"=1 [0]        row kernel uses vectorised LDG instead of the TMA pipeline;
 *  "no_fast_nucleus"=1 [0]  top-p rows skip nucleus_fast_kernel and nucleus_hist_kernel (exact band search only);
 *  "no_hist_nucleus"=1 [0]  flat top-p rows skip nucleus_hist_kernel (band search instead of the histogram select);
 *  "no_tma_nucleus"=0 [1]   top-p rows: first sweep of nucleus_fast_kernel as a launch of the TMA row pipeline (off by
 *                           default: faster on flat rows, slower on LLM-like rows);
 *  "no_fast_ngram"=1 [0]    greedy n-gram verify takes the exact kernels;
 *  "no_fused_tail"=1 [0]    exact_rows + sample_partial kernels instead of tail_fused_kernel;
 *  "no_pdl"=1 [0]           plan / fused tail launched without programmatic dependent launch;
 *  "tma_ngram"=0 [1]        greedy n-gram verify on 16-bit rows: LDG arg-max kernel instead of the TMA row pipeline;
 *  "chunks"=n [2]           batch chunks pipelined on two streams (bf16/fp16, B >= 64 n; 1 = off; "no_overlap"=1 is
 *                           the same as "chunks"=1);  "p1_ctas"=k [3] row-kernel CTAs per SM for chunks > 0;
 *  "tf_ch"=k [20]           CTAs per sequence of the fused tail;  "tf_balance"=0 [1] slices NOT rounded to a multiple of
 *                           8 segments (one per warp);
 *  "tail_slots"=0 [1]       fused tail exchanges through atomics + counters (tail_fused_kernel) instead of
 *                           self-validating words (tail_slots_kernel);
 *  "static_rows"=1 [0]      TMA row kernel: rows assigned by blockIdx instead of claimed from a counter;
 *  "small_b"=B [0]          batches of <= B sequences (plain modes) take the one-launch cluster-per-sequence kernel
 *                           (measured slower than the pipeline: opt-in);  "small_cl"=8 [16] CTAs per cluster;
 *  "no_rowsel"=1 [0]        masked modes skip the streamed selection kernel (rowsel_tma_kernel);
 *  "no_klist"=1 [0]         masked modes draw by a sweep over the row instead of the kept-token lists;
 *  "split_lists"=1 [0]      masked modes: plan and the list draw as two launches (plan_kernel + sample_lists_kernel);
 *  "mega"=1 [0]             plain modes as one persistent cooperative launch (experimental, slower);
 *  "reset"                  every option back to its default. */
SPECDEC_API int specdec_set_option(const char* name, int value);
/* Test hook: copies 16 device-side counters of nucleus_hist_kernel to out16 (host memory; synchronises):
 * [0..6] failed attempts by reason, [7] rows left to the slow path, [8] rows resolved, [9] attempts. */
SPECDEC_API int specdec_debug_stats(unsigned long long* out16, int reset);
/* Tuning hook: with option "mega_dbg"=1 the persistent verify kernel stamps %globaltimer (ns) into its workspace:
 * [0] first CTA start, [1] last row-streaming CTA done, [2] last CTA done, [3] first row-streaming CTA done, then 8 per
 * sequence: plan start, plan published, first / last exact item started, normalisers complete, finalize start / end.
 * Copies min(n, 16 + 8 B) values to host_out (synchronises the device). */
SPECDEC_API int specdec_debug_timeline(const void* workspace, int B, int gamma, int V, unsigned long long* host_out, int n);

/* LogitsProcessor.__call__ materialised: probs[rows,V] fp32 = softmax(_process(logits)/T)
 * (utils/logits_processor.py:13-15).  row_stats (nullable) receives 8 floats per row:
 * {max, S32, inv, cut, jcut, n/a, n/a, n/a}. */
SPECDEC_API int specdec_process_probs(const void* logits, int dtype, int64_t rows, int V, int64_t stride,
                          float temperature, int top_k, float top_p, float* probs, float* row_stats,
                          void* workspace, size_t workspace_bytes, specdec_stream_t stream);

/* processor + sample() fused for AR / drafter steps (sampling/base_decoding.py:51-57,
 * sampling/speculative_decoding.py:120-124): tok[r] ~ processor(logits[r]); ptok[r] = its probability.
 * u [rows] nullable (Philox keyed by seq_id0+r, lane 0x20000+lane_id). */
SPECDEC_API int specdec_sample_rows(const void* logits, int dtype, int64_t rows, int V, int64_t stride,
                        float temperature, int top_k, float top_p, int sample_mode, const float* u,
                        uint64_t philox_seed, uint64_t philox_offset, int64_t seq_id0, int lane_id,
                        int64_t* tok, float* ptok, void* workspace, size_t workspace_bytes,
                        specdec_stream_t stream);

/* LogitsProcessor.sample(probs) on materialised fp32 probabilities [rows,V]
 * (utils/logits_processor.py:35-36 greedy, :48-49 restated as inverse CDF). */
SPECDEC_API int specdec_sample_probs(const float* probs, int64_t rows, int V, int sample_mode, const float* u,
                         int64_t* tok, specdec_stream_t stream);

/* Dumps the exact uniforms specdec_verify would use (so tests can inject them into the oracle). */
SPECDEC_API int specdec_philox_uniform(uint64_t seed, uint64_t offset, int64_t seq_id0, int B, int gamma,
                           float* u_accept, float* u_sample, specdec_stream_t stream);

/*
 * Per-sequence KV rollback on a static cache: prune_cache / prune_tuple_cache
 * (utils/caching.py:6-55; call sites sampling/speculative_decoding.py:163-165) generalised to a
 * different discard count per sequence.  tensor_ptrs: DEVICE array of n_tensors device pointers,
 * each a [B,H,S_max,D] tensor of elem_bytes-wide elements; seq_lens[B] in/out.  The length vector IS the
 * rollback (the valid prefix [0, seq_lens[b]) equals the reference's view): with zero_fill == 0 the tensors are
 * not touched (tensor_ptrs may be NULL, n_tensors 0) and the call is one ~3 us launch; zero_fill != 0 also clears
 * the discarded positions (one CTA per (tensor, sequence), 16-byte stores).
 */
SPECDEC_API int specdec_prune_kv(void* const* tensor_ptrs, int n_tensors, int B, int H, int64_t S_max, int64_t D,
                     int elem_bytes, int32_t* seq_lens, const int32_t* discard, int zero_fill,
                     specdec_stream_t stream);

/*
 * Device n-gram tables: NGramStorage / OneLevelNGramStorage (ngram_assisted/ngram_storage.py:73-249).
 * One logical table per table id (sequence); capacity is per table.  `one_level`!=0 gives
 * OneLevelNGramStorage semantics (single context length n-1).
 * The handle is a host pointer; table memory is allocated once at create (cudaMalloc), mirroring the
 * reference's constructor; create returns after the tables are zeroed (it synchronises the NULL stream).
 * table_ids must lie in [0, n_tables): the kernels index the tables with them unchecked (the Python host validates).
 */
typedef struct specdec_ngram specdec_ngram_t;
SPECDEC_API int specdec_ngram_create(specdec_ngram_t** out, int n, int vocab_size, int n_tables, int grams_per_table,
                         int counts_per_table, int one_level);
SPECDEC_API int specdec_ngram_destroy(specdec_ngram_t* t);
SPECDEC_API int specdec_ngram_reset(specdec_ngram_t* t, specdec_stream_t stream);
/* ids [B, max_len] int64 row-major with per-row valid length lens[B]; table_ids[B] (NULL = all table 0). */
SPECDEC_API int specdec_ngram_initialize(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                             int B, int64_t max_len, specdec_stream_t stream);
/* next_tokens [B, m] int64 */
SPECDEC_API int specdec_ngram_update(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                         int B, int64_t max_len, const int64_t* next_tokens, int m, specdec_stream_t stream);
/* gamma chained next_token() calls (ngram_assisted/ngram_assisted.py:95-99): drafts[B,gamma],
 * known[B,gamma]; unknown positions take fallback[B,gamma] (the reference draws torch.randint). */
SPECDEC_API int specdec_ngram_lookup_chain(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                               int B, int64_t max_len, int gamma, const int64_t* fallback, int64_t* drafts,
                               uint8_t* known, specdec_stream_t stream);
/* has_gram (ngram_storage.py:98-106 / :181-193), exact: out[b] = 1 iff the final token of row b was ever counted after
 * the context made of the row's last j tokens (the reference's context includes that final token), longest j first. */
SPECDEC_API int specdec_ngram_has_gram(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                           int B, int64_t max_len, uint8_t* out, specdec_stream_t stream);
/* seed of the fallback tokens lookup_chain draws ON THE DEVICE for unknown contexts when `fallback` is NULL
 * (Philox keyed by seed, call number, sequence, position; the reference draws torch.randint, ngram_storage.py:84,165) */
SPECDEC_API int specdec_ngram_seed(specdec_ngram_t* t, uint64_t seed);
/* device int32[2]: {overflow flag, entries used (max over tables)} */
SPECDEC_API int specdec_ngram_status(specdec_ngram_t* t, int32_t* host_out2);

/*
 * Decode-loop helpers (csrc/engine.cu).
 *
 * specdec_topk_ids: ids of the k largest logits of each row, value descending / index ascending on ties, -1 when the
 * row has fewer than k elements -- the filler tokens of ngram_assisted/ngram_assisted.py:149-155
 * (`torch.topk(p[..., i, :], filler_top_k)`; probabilities are monotone in the logits, no V-wide probability row is
 * written).  out_ids [rows, k] int64.
 *
 * specdec_batch_writeback: the per-sequence bookkeeping after a batched verify (engine/infer_engine.py:300-336) on
 * device-resident state, one launch, no host read-back: for every sequence with finished[b] == 0
 *   acc = first_stop >= 0 ? first_stop + 1 : n_accepted;  n_acc[b] += acc;
 *   rejected (no stop, n < gamma): generated[b, step + n] = next_token;  generated[b, step + acc + 1 ..] = 0;
 *   finished[b] = stop hit || (rejected && next_token is an end token).
 * step: host value, or read from *step_dev when step_dev != NULL (graph-captured loops).  n_active_out (nullable):
 * device int32 that receives the number of sequences still running.
 */
SPECDEC_API int specdec_topk_ids(const void* logits, int dtype, int64_t rows, int V, int64_t row_stride, int k, int64_t* out_ids,
                     specdec_stream_t stream);
SPECDEC_API int specdec_batch_writeback(int B, int gamma, const int32_t* n_accepted, const int32_t* first_stop,
                            const int64_t* next_token, int64_t* generated, int64_t gen_stride, const int64_t* step_dev,
                            int64_t step, uint8_t* finished, int64_t* n_acc, const int64_t* end_tokens, int n_end,
                            int32_t* n_active_out, specdec_stream_t stream);

/*
 * Multi-GPU: all-gather of the packed per-sequence results by peer-to-peer stores over NVLink / NVSwitch
 * instead of an NCCL kernel.  peer_bufs_dev / peer_flags_dev: DEVICE arrays of `world` pointers to every rank's gather
 * buffer / flag array (peer-mapped, e.g. torch.distributed._symmetric_memory).  specdec_peer_publish copies
 * packed_local[0, n_words) to peer_bufs[p] + dst_off_words for every p, then releases peer_flags[p][flag_off_words]
 * = seq; specdec_peer_wait (one warp) returns on the stream once flags_local[r] >= seq for every r < world
 * (*status = 1 after a bounded wait).  wait_flags_local != NULL makes specdec_peer_publish ALSO perform that wait (for
 * an earlier step's slot, wait_seq) in an extra CTA of the same launch.  Both ride the caller's stream; neither touches
 * the host.
 */
SPECDEC_API int specdec_peer_publish(const int32_t* packed_local, int n_words, void* const* peer_bufs_dev, int64_t dst_off_words,
                         int world, void* const* peer_flags_dev, int64_t flag_off_words, int32_t seq,
                         const int32_t* wait_flags_local, int32_t wait_seq, int32_t* status, specdec_stream_t stream);
SPECDEC_API int specdec_peer_wait(const int32_t* flags_local, int world, int32_t seq, int32_t* status, specdec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
