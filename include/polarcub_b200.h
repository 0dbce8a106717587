/*
 * polarcub_b200.h -- C-ABI of the B200-native batched polar-code engine.
 *
 * Drop-in boundary for the encode / decode hot path of benjilieber/polarcub.  Every entry point takes
 * plain pointers and sizes (no torch / C++ types), returns 0 on success or a negative pc_status, never
 * throws, and launches on the CUDA stream passed as `stream` (a cudaStream_t cast to void*).
 * All `d_*` pointers are DEVICE pointers owned by the caller; `h_*` pointers are host pointers read
 * before the call returns.  A plan is immutable after creation and may be shared between threads.
 *
 * Bit packing: bit i of a frame lives in word i/32, bit position i%32 (LSB first).  Codewords are in
 * the reference's index order (x = u B_N F^{(x)n}, BinaryPolarEncoderDecoder.py:321-323).
 *
 * Reference interface replaced by each entry point (paths relative to the reference root):
 *   pc_plan_create            BinaryPolarEncoderDecoder.__init__ / initializeFrozenOrInformation...
 *                             (BinaryPolarEncoderDecoder.py:16-44), QaryPolarEncoderDecoder.__init__
 *                             (QaryPolarEncoderDecoder.py:27-63)
 *   pc_encode_bits            BinaryPolarEncoderDecoder.encode            (BinaryPolarEncoderDecoder.py:46-69)
 *                             with a uniform prior; polarTransformOfBits is its inverse (:494-516)
 *   pc_sc_decode_probs        BinaryPolarEncoderDecoder.decode            (BinaryPolarEncoderDecoder.py:71-99,
 *                             recursion :223-325, arithmetic BinaryMemorylessVectorDistribution.py:15-87)
 *   pc_sc_decode_symbols      same, fused with makeBinaryMemorylessVectorDistribution(length, yvec)
 *                             (ScalarDistributions/BinaryMemorylessDistribution.py:245-258)
 *   pc_qsc_encode             QaryPolarEncoderDecoder.encode              (QaryPolarEncoderDecoder.py:65-88)
 *   pc_qsc_decode_probs       QaryPolarEncoderDecoder.decode              (QaryPolarEncoderDecoder.py:90-116,
 *                             recursion :318-401, arithmetic QaryMemorylessVectorDistribution.py:26-118)
 *   pc_scl_decode_probs       QaryPolarEncoderDecoder.listDecode          (QaryPolarEncoderDecoder.py:118-227,
 *                             recursion :403-757, helpers :759-820, :867-872)
 *   pc_scl_decode_packed      the same entry point for q = 2 on bit-packed buffers, float64 pairs or, fused with
 *   pc_scl_decode_symbols     makeQaryMemorylessVectorDistribution(q, length, yvec) (ScalarDistributions/
 *                             QaryMemorylessDistribution.py:757-776), uint8 channel output symbols + the channel table
 *   pc_trellis_decode         BinaryPolarEncoderDecoder.decode over CollectionOfBinaryTrellises (uniform prior), fused with
 *                             buildCollectionOfBinaryTrellises_uniformInput_deletion's per-sub-word trellis construction
 *                             (VectorDistributions/BinaryTrellis.py:206-438, CollectionOfBinaryTrellises.py:55-129)
 */
#ifndef POLARCUB_B200_H
#define POLARCUB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pc_plan pc_plan;

enum pc_status {
    PC_OK = 0,
    PC_ERR_INVALID = -1,   /* bad argument (null pointer, unsupported q / n / L, ...) */
    PC_ERR_NOMEM = -2,     /* host or device allocation failed, or the workspace is too small */
    PC_ERR_CUDA = -3,      /* a CUDA runtime call failed; see pc_last_error() */
    PC_ERR_UNSUPPORTED = -4
};

/* library / build identification */
int pc_version(void);
/* thread-local text of the last error returned on this thread ("" if none) */
const char *pc_last_error(void);
/* number of kernels launched by this library since load (all threads); used by bench.py's gpu_launches */
unsigned long long pc_kernel_launch_count(void);

/* ---- plans -------------------------------------------------------------------------------------- */
/* q: alphabet size (2 = binary path).  n: log2 of the block length, 0 <= n <= 24.
 * h_frozen_mask[N]: 1 = frozen u index.  h_frozen_vals[N]: the value every frozen u_i takes
 * (binary: (0.5 >= r_i) ? 0 : 1 for a uniform prior, BinaryPolarEncoderDecoder.py:258-262; q-ary: 0,
 * QaryPolarEncoderDecoder.py:351); entries at information positions are ignored.
 * The plan is bound to the CUDA device current at creation. */
int pc_plan_create(int q, int n, const uint8_t *h_frozen_mask, const uint8_t *h_frozen_vals, pc_plan **out);
void pc_plan_destroy(pc_plan *plan);
int pc_plan_k(const pc_plan *plan);      /* number of information symbols */
int pc_plan_length(const pc_plan *plan); /* N */
/* number of schedule nodes after rate-0 pruning (diagnostic) */
int pc_plan_schedule_len(const pc_plan *plan);

/* ---- binary encode ------------------------------------------------------------------------------ */
/* d_info_packed [B][ceil(k/32)] -> d_cw_packed [B][ceil(N/32)].  Frozen bits come from the plan. */
int pc_encode_bits(const pc_plan *plan, const uint32_t *d_info_packed, uint32_t *d_cw_packed, int64_t B,
                   void *stream);
/* inverse map x -> u (polarTransformOfBits): d_cw_packed [B][ceil(N/32)] -> d_u_packed [B][ceil(N/32)] */
int pc_polar_transform_bits(int n, const uint32_t *d_cw_packed, uint32_t *d_u_packed, int64_t B, void *stream);

/* ---- binary SC decode --------------------------------------------------------------------------- */
#define PC_INPUT_SYMBOLS 0
#define PC_INPUT_PROBS 1
/* bytes of device scratch the decoders want for a batch of B frames (they accept less and then work in
 * smaller chunks, down to 32 frames; PC_ERR_NOMEM below that) */
size_t pc_sc_workspace_bytes(const pc_plan *plan, int64_t B, int input_kind);
/* the same for pc_sc_decode_symbols when the channel table is known: erasure-type tables (every row hard knowledge or an
 * exact erasure, e.g. makeBEC, BinaryMemorylessDistribution.py:493-499) let the large-block decoder keep one byte per
 * tree element instead of a float64, so it wants less scratch per frame and takes larger batches */
size_t pc_sc_workspace_bytes_symbols(const pc_plan *plan, int64_t B, const double *h_table, int Y);
/* frames one launch of the binary SC decoder keeps resident on the device (all SMs busy); batches and host-pipeline chunks
 * are best sized in whole multiples of it.  No reference counterpart (the reference decodes one frame per call). */
int64_t pc_sc_wave_frames(const pc_plan *plan);

/* d_xy [B][N][2] float64: entry i holds P(X=0,Y=y_i), P(X=1,Y=y_i) exactly as the reference's
 * xyVectorDistribution.probs.  Outputs: d_cw_packed [B][ceil(N/32)] (re-encoded codeword) and
 * d_info_packed [B][ceil(k/32)].  Arithmetic is float64 and BIT-IDENTICAL to the reference (same
 * products, sums, max-normalisation and tie rule). */
int pc_sc_decode_probs(const pc_plan *plan, const double *d_xy, int64_t B, uint32_t *d_cw_packed,
                       uint32_t *d_info_packed, void *d_workspace, size_t workspace_bytes, void *stream);

/* Genie pass (BinaryPolarEncoderDecoder.genieSingleDecodeSimulatioan, BinaryPolarEncoderDecoder.py:114-178): every index is
 * frozen to a known bit -- d_u_packed [B][ceil(N/32)], the u vector of each frame, bit i = u_i -- and the decoder captures
 * marginalizedUProbs (:268-273): d_marg [B][N][2] float64 = P(U_i = x | u_0^{i-1}, y), normalised as
 * calcMarginalizedProbabilities does.  d_cw_packed receives the codeword of u.  The plan's frozen set is ignored (only its
 * length is used).  2 <= N <= 65536. */
size_t pc_sc_genie_workspace_bytes(const pc_plan *plan, int64_t B);
int pc_sc_genie_probs(const pc_plan *plan, const double *d_xy, const uint32_t *d_u_packed, int64_t B, uint32_t *d_cw_packed,
                      double *d_marg, void *d_workspace, size_t workspace_bytes, void *stream);

/* Non-uniform a-priori distributions (frozen bits depend on the data, BinaryPolarEncoderDecoder.py:258-262).
 * d_rnd: randomlyGeneratedNumbers (:33-44), float64, row f at d_rnd + f * rnd_row_stride (0: one vector of N shared by all
 * rows; genie batches use one row per frame).  Workspace: pc_sc_genie_workspace_bytes(plan, rows).
 * pc_sc_decode_probs_prior -- decode (:71-99 with both trees, :277-317): d_pairs [rows][N][2] with rows = 2 B: row 2f holds
 *   frame f's xyVectorDistribution.probs, row 2f+1 its xVectorDistribution.probs (the two trees run in adjacent lanes).
 *   Outputs have `rows` rows too (d_cw_packed [rows][ceil(N/32)], d_info_packed [rows][ceil(k/32)]); rows 2f and 2f+1 are equal.
 *   d_marg (optional) [rows][N][2]: the captured leaf probabilities of both trees.
 * pc_sc_encode_prior -- encode (:46-69): d_x [B][N][2] the a-priori probabilities, d_u_packed [B][ceil(N/32)] the information
 *   bits placed at their u positions (frozen positions ignored); frozen u_i = 0 iff P(u_i = 0 | past) >= r_i.  d_marg
 *   (optional) [B][N][2] captures P(u_i | past) (genieSingleEncodeSimulatioan, :180-221). */
int pc_sc_decode_probs_prior(const pc_plan *plan, const double *d_pairs, const double *d_rnd, int64_t rnd_row_stride, int64_t rows,
                             uint32_t *d_cw_packed, uint32_t *d_info_packed, double *d_marg, void *d_workspace,
                             size_t workspace_bytes, void *stream);
int pc_sc_encode_prior(const pc_plan *plan, const double *d_x, const uint32_t *d_u_packed, const double *d_rnd,
                       int64_t rnd_row_stride, int64_t B,
                       uint32_t *d_cw_packed, double *d_marg, void *d_workspace, size_t workspace_bytes, void *stream);

/* d_y [B][N] uint8 channel output symbols; h_table [Y][2] float64 = the channel's joint probabilities
 * (BinaryMemorylessDistribution.probs), 1 <= Y <= 16. */
int pc_sc_decode_symbols(const pc_plan *plan, const uint8_t *d_y, int64_t B, const double *h_table, int Y,
                         uint32_t *d_cw_packed, uint32_t *d_info_packed, void *d_workspace, size_t workspace_bytes,
                         void *stream);

/* ---- q-ary SC ------------------------------------------------------------------------------------ */
/* symbols are uint8 in [0, q).  d_info [B][k] -> d_cw [B][N] */
int pc_qsc_encode(const pc_plan *plan, const uint8_t *d_info, uint8_t *d_cw, int64_t B, void *stream);
size_t pc_qsc_workspace_bytes(const pc_plan *plan, int64_t B);
int64_t pc_qsc_wave_frames(const pc_plan *plan); /* see pc_sc_wave_frames */
/* d_xy [B][N][q] float64 (linear domain).  Outputs d_cw [B][N], d_info [B][k] uint8. */
int pc_qsc_decode_probs(const pc_plan *plan, const double *d_xy, int64_t B, uint8_t *d_cw, uint8_t *d_info,
                        void *d_workspace, size_t workspace_bytes, void *stream);

/* d_y [B][N] uint8 channel output symbols, h_table [Y][q] float64 = QaryMemorylessDistribution.probs (rows: output symbols),
 * 1 <= Y <= 16: pc_qsc_decode_probs fused with makeQaryMemorylessVectorDistribution(length, yvec)
 * (ScalarDistributions/QaryMemorylessDistribution.py:757-766) -- q x 8 times less input traffic. */
int pc_qsc_decode_symbols(const pc_plan *plan, const uint8_t *d_y, int64_t B, const double *h_table, int Y, uint8_t *d_cw,
                          uint8_t *d_info, void *d_workspace, size_t workspace_bytes, void *stream);

/* The LOG DOMAIN of the q-ary SC decoder: QaryPolarEncoderDecoder(q, length, frozenSet, seed, use_log=True).decode
 * (QaryPolarEncoderDecoder.py:27,47; the `if self.use_log` branches of QaryMemorylessVectorDistribution,
 * VectorDistributions/QaryMemorylessVectorDistribution.py:31-42 logaddexp convolution, :50-62, :74-82, :97-118 logsumexp
 * normalisation).  d_xy_log [B][N][q] float64 natural logarithms (-inf for probability 0); h_log_table [Y][q] likewise
 * (makeQaryMemorylessVectorDistribution(..., use_log=True), QaryMemorylessDistribution.py:757-776).  The device's exp / log1p
 * replace the host libm's: decisions equal the reference's except at ties within ~1e-15 of the log values. */
int pc_qsc_decode_logprobs(const pc_plan *plan, const double *d_xy_log, int64_t B, uint8_t *d_cw, uint8_t *d_info,
                           void *d_workspace, size_t workspace_bytes, void *stream);
int pc_qsc_decode_symbols_log(const pc_plan *plan, const uint8_t *d_y, int64_t B, const double *h_log_table, int Y,
                              uint8_t *d_cw, uint8_t *d_info, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- SC-list decoding (any q in {2,3,4,5}; binary SCL is q = 2) --------------------------------------- */
/* QaryPolarEncoderDecoder.listDecode with actualInformation (genie selection, the form ir() uses,
 * QaryPolarEncoderDecoder.py:856).  d_xy [B][N][q] float64 (linear domain); d_frozen_values [B][N-k] uint8 (the
 * explicit frozen values, consumed in u order like frozenValuesIterator); d_actual_info [B][k] uint8.
 * Outputs: d_info [B][k] uint8 (the word listDecode returns) and d_prob_result [B] int32 (ProbResult value,
 * QaryPolarEncoderDecoder.py:18-24).  Optional (all null, or the first three non-null): the final list --
 * d_list_size [B] int32, d_list_prob [B][L] float64 (normalised metrics, list order), d_actual_prob [B] float64,
 * d_list_info [B][L][k] uint8 (may be null on its own).  1 <= L <= 32 (q <= 3) or L <= 8 (q = 4, 5); N >= 2. */
size_t pc_scl_workspace_bytes(const pc_plan *plan, int L, int64_t B, int want_list);
/* frames one launch of the list decoder keeps resident on the device (see pc_sc_wave_frames) */
int64_t pc_scl_wave_frames(const pc_plan *plan, int L);
int pc_scl_decode_probs(const pc_plan *plan, int L, const double *d_xy, const uint8_t *d_frozen_values,
                        const uint8_t *d_actual_info, int64_t B, uint8_t *d_info, int32_t *d_prob_result,
                        int32_t *d_list_size, double *d_list_prob, double *d_actual_prob, uint8_t *d_list_info,
                        void *d_workspace, size_t workspace_bytes, void *stream);

/* listDecode in the LOG DOMAIN (use_log=True: QaryPolarEncoderDecoder.py:140, :176, :232, :436-566, :594-663, :765, :784,
 * :813, :869): d_xy_log [B][N][q] natural logarithms; d_list_prob / d_actual_prob are log metrics (list maximum 0).  Same
 * buffers otherwise; q in {2,3,4,5} (q = 2 runs on the generic frame-per-lane kernel here).  Metrics agree with the
 * reference's to ~1e-14 (device exp / log1p), decisions on tie-free inputs are identical. */
size_t pc_scl_workspace_bytes_log(const pc_plan *plan, int L, int64_t B, int want_list);
int pc_scl_decode_logprobs(const pc_plan *plan, int L, const double *d_xy_log, const uint8_t *d_frozen_values,
                           const uint8_t *d_actual_info, int64_t B, uint8_t *d_info, int32_t *d_prob_result,
                           int32_t *d_list_size, double *d_list_prob, double *d_actual_prob, uint8_t *d_list_info,
                           void *d_workspace, size_t workspace_bytes, void *stream);

/* Binary (q = 2) listDecode on BIT-PACKED buffers (bit i of a row in word i/32, position i%32), 2 <= N <= 8192.
 * Channel input, exactly one of:
 *   d_xy [B][N][2] float64 = xyVectorDistribution.probs (linear domain), or
 *   d_y  [B][N] uint8 channel output symbols, every y < Y, with h_table [Y][2] float64 = the rows
 *        makeQaryMemorylessVectorDistribution would copy into probs (QaryMemorylessDistribution.py:757-776), 1 <= Y <= 256:
 *        16 x less input traffic; results are those of pc_scl_decode_probs on h_table[y].
 * d_frozen_packed [B][ceil((N-k)/32)]: the explicit frozen values in u order (frozenValuesIterator), or NULL = all zero.
 * d_actual_info_packed [B][ceil(k/32)].  Outputs: d_info_packed [B][ceil(k/32)], d_prob_result [B] int32 (ProbResult).
 * Optional final list (all null, or the first three non-null): d_list_size [B], d_list_prob [B][L], d_actual_prob [B],
 * d_list_info_packed [B][L][ceil(k/32)] (may be null on its own).  Workspace: pc_scl_workspace_bytes_packed. */
size_t pc_scl_workspace_bytes_packed(const pc_plan *plan, int L, int64_t B, int want_list);
int pc_scl_decode_packed(const pc_plan *plan, int L, const double *d_xy, const uint8_t *d_y, const double *h_table, int Y,
                         const uint32_t *d_frozen_packed, const uint32_t *d_actual_info_packed, int64_t B,
                         uint32_t *d_info_packed, int32_t *d_prob_result, int32_t *d_list_size, double *d_list_prob,
                         double *d_actual_prob, uint32_t *d_list_info_packed, void *d_workspace, size_t workspace_bytes,
                         void *stream);
/* pc_scl_decode_packed on channel symbols without the list outputs */
int pc_scl_decode_symbols(const pc_plan *plan, int L, const uint8_t *d_y, const double *h_table, int Y,
                          const uint32_t *d_frozen_packed, const uint32_t *d_actual_info_packed, int64_t B,
                          uint32_t *d_info_packed, int32_t *d_prob_result, void *d_workspace, size_t workspace_bytes,
                          void *stream);

/* ---- channel simulation and guard-band plumbing on the device (the steps either side of the decoders) ------------------- */
/* Discrete memoryless channel: y[f][i] = the first output symbol whose cumulated P(y | x[f][i]) reaches a uniform draw -- the
 * inverse-CDF walk of the reference's simulators (test3.py:35-54, test2.py:29-65).  h_cond [X][Y] float64 rows P(. | x).
 * Input symbols: d_x [B][N] uint8, or (binary) d_x_packed [B][ceil(N/32)] -- exactly one non-null.  The generator is
 * counter-based, keyed by (seed, frame0 + f, i): the output of a frame does not depend on how the batch is split over calls or
 * ranks.  Workspace: X * Y * 8 bytes. */
int pc_channel_simulate_dmc(const uint8_t *d_x, const uint32_t *d_x_packed, int64_t B, int N, int X, int Y, const double *h_cond,
                            uint64_t seed, int64_t frame0, uint8_t *d_y, void *d_workspace, size_t workspace_bytes, void *stream);
/* BI-AWGN: y = (1 - 2 x) + sigma * N(0, 1); d_y_real [B][N] float64 and / or d_y_quantised [B][N] uint8 =
 * clamp(floor((y + ymax) / (2 ymax / Y)), 0, Y - 1), the symbol input of pc_sc_decode_symbols / pc_scl_decode_symbols. */
int pc_channel_simulate_biawgn(const uint8_t *d_x, const uint32_t *d_x_packed, int64_t B, int N, double sigma, uint64_t seed,
                               int64_t frame0, int Y, double ymax, uint8_t *d_y_quantised, double *d_y_real, void *stream);
/* Channel symbols of small alphabets travel packed between host and device (a BSC output is one bit, a BEC output two):
 * `bits` in {1, 2, 4} bits per symbol, symbol i of the stream in bits [i * bits, (i + 1) * bits) of the little-endian word stream;
 * `count` symbols, count * bits a multiple of 32.  The decoders take the unpacked uint8 symbols (pc_sc_decode_symbols, ...). */
int pc_unpack_symbols(const void *d_packed, int bits, int64_t count, uint8_t *d_symbols, void *stream);
int pc_pack_symbols(const uint8_t *d_symbols, int bits, int64_t count, void *d_packed, void *stream);
/* Guardbands.addDeletionGuardBands (Guardbands.py:4-44) on a batch: d_encoded [B][2^n] uint8 -> d_out [B][pc_guard_band_length]
 * (zeros between the halves of every block above level n0, `ones` ones around each sub-word).  Workspace: 4 * 2^(n-n0) bytes. */
int pc_guard_band_length(int n, int n0, double xi, int ones);
int pc_add_guard_bands(const uint8_t *d_encoded, int64_t B, int n, int n0, double xi, int ones, uint8_t *d_out, void *d_workspace,
                       size_t workspace_bytes, void *stream);
/* deletionChannelSimulation (VectorDistributions/BinaryTrellis.py:441-461): symbol i of d_in [B][len] survives iff its uniform
 * draw is >= deletion_prob; survivors in order at d_out [B][len] (zero padded), their number in d_out_len [B]. */
int pc_deletion_channel(const uint8_t *d_in, int64_t B, int len, double deletion_prob, uint64_t seed, int64_t frame0, uint8_t *d_out,
                        int32_t *d_out_len, void *stream);
/* Guardbands.removeDeletionGuardBands (Guardbands.py:47-93): d_received [B][stride] uint8 with lengths d_received_len [B] (null:
 * every word has `stride` symbols) -> the 2^(n-n0) trimmed sub-words d_sub_bits [B][2^(n-n0)][maxlen], d_sub_len [B][2^(n-n0)]
 * (the inputs of pc_trellis_decode); *d_overflow is set when a sub-word is longer than maxlen. */
int pc_remove_guard_bands(const uint8_t *d_received, const int32_t *d_received_len, int64_t B, int stride, int n, int n0, int maxlen,
                          uint8_t *d_sub_bits, int32_t *d_sub_len, int32_t *d_overflow, void *stream);

/* ---- deletion channel: SC decoding over a collection of trellises ---------------------------------------- */
/* The received word is split by the caller into T = 2^(n-n0) trimmed sub-words (Guardbands.removeDeletionGuardBands,
 * Guardbands.py:47-63): d_sub_bits [B][T][maxlen] uint8 (0/1), d_sub_len [B][T] int32 (lengths <= maxlen).  Each sub-word
 * becomes a trellis of 2^n0 inputs exactly as buildTrellis_uniformInput_deletion(subword, 2^n0, deletion_prob, True,
 * ones) builds it; the first n0 decoding levels transform trellises, the rest is memoryless SC decoding.
 * Outputs as pc_sc_decode_probs.  d_first_collapse (optional, may be null) [B][T][2] float64 receives the unnormalised
 * pairs of the first collapsed vector (all-minus descent), for parity checks.  1 <= n0 <= min(4, n), ones <= 16,
 * maxlen <= 250. */
size_t pc_trellis_workspace_bytes(const pc_plan *plan, int n0, int maxlen, int64_t B);
int pc_trellis_decode(const pc_plan *plan, int n0, double deletion_prob, int ones, const uint8_t *d_sub_bits,
                      const int32_t *d_sub_len, int maxlen, int64_t B, uint32_t *d_cw_packed, uint32_t *d_info_packed,
                      double *d_first_collapse, void *d_workspace, size_t workspace_bytes, void *stream);

/* genie pass over trellis collections (see pc_sc_genie_probs); workspace: pc_trellis_workspace_bytes; n - n0 >= 1 */
int pc_trellis_genie(const pc_plan *plan, int n0, double deletion_prob, int ones, const uint8_t *d_sub_bits,
                     const int32_t *d_sub_len, int maxlen, const uint32_t *d_u_packed, int64_t B, uint32_t *d_cw_packed,
                     double *d_marg, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- code construction (HOST only: no device buffers, no stream) -------------------------------------- */
/* Tal-Vardy degrading construction for a binary-input memoryless channel with a uniform input: h_pe[i], i < 2^n, is the
 * error probability of the degraded synthetic channel i (MSB-first minus / plus order) -- the Pevec of
 * calcFrozenSet_degradingUpgrading(n, L, eps, None, xyDistribution) (ScalarDistributions/BinaryMemorylessDistribution.py:
 * 620-680: minusTransform().degrade(L) / plusTransform().degrade(L) per level, errorProb() per leaf), float64-identical to
 * the reference.  h_table [Y][2] = xyDistribution.probs; `threads` host threads share the nodes of a level. */
int pc_tv_degrade_pe(int n, int L, const double *h_table, int Y, double *h_pe, int threads);
/* The same for a q-ary input alphabet: the Pevec of QaryMemorylessDistribution's calcTVAndPe_degradingUpgrading(n, L, None,
 * xyDistribution) (ScalarDistributions/QaryMemorylessDistribution.py:934-990; degrade = degrade_dynamic :215-260 over the
 * q-1 one-hot binary channels :98-153).  h_table [Y][q] = xyDistribution.probs. */
int pc_tv_degrade_pe_qary(int q, int n, int L, const double *h_table, int Y, double *h_pe, int threads);

/* ---- Monte-Carlo counters and measurement hooks ---------------------------------------------------- */
/* d_out3[0..2] += {B, frames whose first nbits differ, differing bits} over packed rows of ceil(nbits/32) words.
 * Replaces the serial comparison loop BinaryPolarEncoderDecoder.py:374-385 / QaryPolarEncoderDecoder.py:907-909;
 * the caller all-reduces the three counters across ranks (NCCL). */
int pc_count_errors(const uint32_t *d_a, const uint32_t *d_b, int64_t B, int nbits, unsigned long long *d_out3,
                    void *stream);
/* When enabled, the library records CUDA events on the launching stream around every launch of the dominant
 * decode kernel; pc_profile_read returns the time (ms) during which at least one of them was running (the union of their
 * intervals: their summed duration when they follow each other on one stream) and the number of launches. */
int pc_profile_enable(int on);
int pc_profile_read(double *total_ms, unsigned long long *launches);

#ifdef __cplusplus
}
#endif
#endif /* POLARCUB_B200_H */
