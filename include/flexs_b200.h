/*
 * flexs_b200 — C ABI of the B200-native virtual-screen hot path of FLEXS.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy / Python types.
 * The reference is pure Python, so the binding a maintainer adds is a ctypes stub
 * (INTEGRATION.md shows it); flexs_b200/_native.py is that stub for this repo.
 *
 * Every entry point below names the reference interface it replaces (paths relative to
 * the reference checkout, samsinai/FLEXS @ dd40916).
 *
 * Conventions
 *   - Return value: 0 on success, a negative FLEXS_E* code on failure;
 *     flexs_last_error() returns a thread-local message for the last failure.
 *   - "d_" pointers are device pointers on the model's device; "h_" pointers are host
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Device entry points enqueue work on `stream` and return without synchronising — except for
 *     the one-time preparation after a weight change or a larger batch than seen before (operand
 *     re-layout, table build, workspace growth), which may allocate and synchronise `stream`.
 *   - Sequences are residue-index arrays uint8[n, seq_len], row-major: the value is
 *     alphabet.index(ch) (flexs/utils/sequence_utils.py:44-47).  The float one-hot of the
 *     reference is never materialised.
 *   - Weights cross the boundary in Keras get_weights() order and layout, fp32:
 *       CNN (flexs/baselines/models/cnn.py:23-54), 12 arrays per member:
 *         W1 (k,A,F) b1 (F) | W2 (k,F,F) b2 | W3 (A-1,F,F) b3 | Wd1 (F,H) bd1 | Wd2 (H,H) bd2 |
 *         Wd3 (H,1) bd3 (1)
 *       MLP (flexs/baselines/models/mlp.py:21-31), 8 arrays per member:
 *         W1 (L*A,H) b1 | W2 (H,H) b2 | W3 (H,H) b3 | W4 (H,1) b4 (1)
 *   - There is no CPU fallback anywhere behind this header: without a CUDA device every
 *     compute entry point fails with FLEXS_ECUDA.
 */
#ifndef FLEXS_B200_H
#define FLEXS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLEXS_B200_ABI_VERSION 1

enum {
    FLEXS_OK = 0,
    FLEXS_EINVAL = -1,   /* bad argument (shape, null pointer, unsupported size)          */
    FLEXS_ECUDA = -2,    /* CUDA runtime error (message in flexs_last_error)              */
    FLEXS_EALPHABET = -3 /* a character outside the alphabet (reference: ValueError)      */
};

enum { FLEXS_KIND_CNN = 1, FLEXS_KIND_MLP = 2 };

/* Kernel variant selection for flexs_model_set_variant (diagnostics / A-B measurements;
 * AUTO is what the product uses). */
enum {
    FLEXS_VARIANT_AUTO = 0,
    FLEXS_VARIANT_SIMPLE = 1, /* one CTA per sequence, any shape                          */
    FLEXS_VARIANT_TILED = 2,  /* register-tiled FP32 FFMA, F == 32                        */
    FLEXS_VARIANT_UMMA = 3,   /* tcgen05 (fp16 hi/lo split) convs, F == 32                */
    FLEXS_VARIANT_UMMA_LUT = 4, /* A == 4: conv1+conv2 as an L2-resident table over 9 residues,
                                 * conv3 + dense head on tcgen05; AUTO uses it for large batches */
    FLEXS_VARIANT_ENUM = 5     /* A^L <= 2^20 (any model kind): the model is evaluated once on all A^L
                                 * sequences by the kernels above, a batch is then one gather per
                                 * candidate; AUTO uses it for batches of at least A^L sequences      */
};

typedef struct flexs_model flexs_model_t;
typedef struct flexs_vae flexs_vae_t;

int flexs_abi_version(void);
const char *flexs_last_error(void);

/* Number of CUDA devices visible, or a negative error. */
int flexs_device_count(void);

/* ---- surrogate objects ---------------------------------------------------------------
 * Replaces the construction of the Keras Sequential in cnn.py:23-54 / mlp.py:21-31.
 * `n_members` > 1 builds the flexs.Ensemble of identical architectures
 * (flexs/ensemble.py:21-40) whose default mean (ensemble.py:24) is applied in the
 * kernel epilogue.  Weights start as zeros; call flexs_model_set_weights.             */
int flexs_cnn_create(int device, int seq_len, int alphabet_size, int num_filters,
                     int hidden_size, int kernel_size, int n_members, flexs_model_t **out);
int flexs_mlp_create(int device, int seq_len, int alphabet_size, int hidden_size,
                     int n_members, flexs_model_t **out);
void flexs_model_destroy(flexs_model_t *m);

/* Number of weight arrays per member (12 or 8) and the element count of array `i`.     */
int flexs_model_num_arrays(const flexs_model_t *m);
int64_t flexs_model_array_size(const flexs_model_t *m, int i);

/* keras Model.set_weights / get_weights for one member (used by cbas_dbas.py:130-144 style
 * cloning and by tools/export_keras_golden.py parity checks).  Host pointers, fp32.       */
int flexs_model_set_weights(flexs_model_t *m, int member, const float *const *h_arrays);
int flexs_model_get_weights(flexs_model_t *m, int member, float *const *h_arrays);

int flexs_model_set_variant(flexs_model_t *m, int variant);
/* Variant the next forward of `n` sequences will run (one of FLEXS_VARIANT_*).          */
int flexs_model_active_variant(const flexs_model_t *m, int64_t n);
/* Kernel launches issued by this model since creation (bench.py's gpu_launches claim).   */
int64_t flexs_model_launch_count(const flexs_model_t *m);

/* ---- K0: encode ----------------------------------------------------------------------
 * Replaces string_to_one_hot (sequence_utils.py:32-47) as called per sequence at
 * keras_model.py:53-56 and :70-73.  d_chars: n*seq_len raw bytes (one byte per residue
 * character).  alphabet: `alphabet_size` bytes.  d_idx receives alphabet.index(ch).
 * d_status (int64[2], device): [0] = number of bytes not in the alphabet, [1] = smallest
 * flat position of such a byte (INT64_MAX if none).  The caller turns [0] != 0 into the
 * reference's ValueError.                                                                */
int flexs_encode_dev(const uint8_t *d_chars, int64_t n_bytes, const char *alphabet,
                     int alphabet_size, uint8_t *d_idx, int64_t *d_status, void *stream);

/* ---- K1/K2: forward ------------------------------------------------------------------
 * Replaces KerasModel._fitness_function (keras_model.py:69-79) including squeeze(axis=1)
 * and np.nan_to_num, and Ensemble._fitness_function's stack+mean (ensemble.py:54-59).
 * d_idx uint8[n, seq_len] -> d_out float32[n].                                           */
int flexs_model_forward_dev(flexs_model_t *m, const uint8_t *d_idx, int64_t n, float *d_out,
                            void *stream);

/* Same call with HOST buffers, the shape Landscape.get_fitness (landscape.py:29-45) has:
 * h_chars = n*seq_len residue characters, h_out = n scores.  Copies are chunked and
 * double-buffered over two streams inside the call; returns after h_out is complete.
 * On FLEXS_EALPHABET *bad_pos (if non-null) is the flat offset of the first bad byte.    */
int flexs_model_score_host(flexs_model_t *m, const char *h_chars, int64_t n,
                           const char *alphabet, float *h_out, int64_t *bad_pos);

/* ---- packed residues: the wire format of the host boundary --------------------------------
 * The float one-hot the reference ships to TensorFlow (keras_model.py:70-75: 4*L*A bytes per
 * sequence) is replaced on the wire by ceil(log2 A) bits per residue: residue i of a row occupies
 * bits [i*b, (i+1)*b) of the row's little-endian bit stream (bit k = bit k%8 of byte k/8), rows
 * padded to whole bytes — 25 bytes for a DNA 100-mer, 149 for a GFP 237-mer.
 * flexs_model_score_host_packed is flexs_model_score_host for rows already in that format (the
 * host packers: flexs_b200/csrc/packstr.c pack_bits for str lists, sequence_utils.pack_indices for
 * arrays); a value >= alphabet_size in a row is reported as FLEXS_EALPHABET with *bad_pos = its
 * flat residue position.  flexs_unpack_dev / flexs_pack_dev convert on the device; d_status as in
 * flexs_encode_dev.                                                                           */
int flexs_bits_per_residue(int alphabet_size);
int64_t flexs_packed_row_bytes(int seq_len, int alphabet_size);
int flexs_unpack_dev(const uint8_t *d_packed, int64_t n, int seq_len, int alphabet_size,
                     uint8_t *d_idx, int64_t *d_status, void *stream);
int flexs_pack_dev(const uint8_t *d_idx, int64_t n, int seq_len, int alphabet_size,
                   uint8_t *d_packed, void *stream);
int flexs_model_score_host_packed(flexs_model_t *m, const uint8_t *h_packed, int64_t n,
                                  float *h_out, int64_t *bad_pos);

/* ---- K3: selection -------------------------------------------------------------------
 * Replaces the final ranking of propose_sequences: np.argsort(preds)[: -B : -1]
 * (adalead.py:171-175, cbas_dbas.py:197-201, cmaes.py:117-122) and
 * np.argsort(preds)[::-1][:B] (dyna_ppo.py:315-319).  Returns the k largest scores in
 * descending order with their indices (+index_offset, for a shard of a larger batch);
 * ties are broken by the LOWER index first (documented deviation: numpy's introsort
 * leaves tie order unspecified).  NaN never appears (forward applies nan_to_num).
 * The reported index of position p is d_index_map[p] when d_index_map is non-NULL (used to
 * merge the all-gathered per-shard top-k lists of a multi-GPU screen), else p + index_offset.
 * If n < k the tail is filled with (-inf, -1).  k <= 4096, n < 2^32.
 * d_work: at least flexs_topk_workspace_bytes(n, k) bytes.                               */
int64_t flexs_topk_workspace_bytes(int64_t n, int k);
int flexs_topk_dev(const float *d_scores, int64_t n, int k, int64_t index_offset,
                   const int64_t *d_index_map, float *d_top_scores, int64_t *d_top_idx,
                   void *d_work, void *stream);

/* ---- K3b: de-duplication before the ranking ----------------------------------------------
 * The reference ranks the keys of a dict (adalead.py:157, cmaes.py:112-115, dyna_ppo.py:310-314): a
 * sequence proposed twice competes once.  d_scores_out[i] = d_scores[i] if row i of d_idx
 * (uint8[n, seq_len]) is the LOWEST-indexed row with that content, else -inf (forward never produces
 * -inf: nan_to_num clamps to +-FLT_MAX), so flexs_topk_dev on d_scores_out returns distinct
 * sequences; winners whose score is -inf mean "fewer than k distinct candidates".  Exact (an
 * open-addressing table keyed by the row bytes).  d_scores_out may alias d_scores.  n < 2^31.
 * d_work: at least flexs_dedup_workspace_bytes(n) bytes.                                        */
int64_t flexs_dedup_workspace_bytes(int64_t n);
int flexs_dedup_scores_dev(const uint8_t *d_idx, int64_t n, int seq_len, const float *d_scores,
                           float *d_scores_out, void *d_work, void *stream);

/* The same table as a lookup: d_rep[i] = the lowest row index whose content equals row i (d_rep[i] == i
 * for a first occurrence).  With the rows of a cache in front of a batch, d_rep answers "already seen?
 * and where is its value?" for every candidate at once — the `seq in seen` / `seq in measured` dict
 * lookups of cmaes.py:85-90 and dyna_ppo.py:310-314.  Same workspace as flexs_dedup_scores_dev.   */
int flexs_dedup_representatives_dev(const uint8_t *d_idx, int64_t n, int seq_len, int64_t *d_rep,
                                    void *d_work, void *stream);

/* ---- K3c: single-launch selection and the multi-GPU merge -----------------------------------
 * flexs_topk_select_dev is flexs_topk_dev (+ flexs_dedup_scores_dev when `unique`) as ONE
 * cooperative kernel launch: radix select with 12-bit digits that stops as soon as the candidates
 * fit a one-CTA sort, then — for `unique` — de-duplication of the best rows only, in rank order
 * (equal rows must carry equal scores, which a deterministic surrogate guarantees).  Same result
 * as the two-call path, bit for bit; when the best max(k, 4096) rows do not hold k distinct
 * sequences although the batch has more rows, d_status[0] = 1 and the outputs are incomplete: the
 * caller then runs flexs_dedup_scores_dev + flexs_topk_dev.  d_status is int32[8] (may be NULL when
 * !unique): [1..4] are diagnostics (radix levels used, candidates sorted, ns selecting, ns in the
 * final one-CTA stage).
 * d_rows uint8[n, row_len] (may be NULL when !unique and d_top_rows is NULL); d_top_rows (may be
 * NULL) receives the winners' rows, uint8[k, row_len].  d_work: flexs_topk_select_workspace_bytes().
 *
 * A rank's message of a sharded screen (screen.py; adalead.py:171-175 over G shards) is
 *   [k] int64 global index | [k] float32 score | [k][seq_len] uint8 rows, padded to 16 bytes
 * (flexs_screen_message_bytes); the three d_top_* pointers of flexs_topk_select_dev can point
 * straight into it.  flexs_screen_merge_dev takes the all-gathered messages of `world` ranks
 * (rank-major) and returns the global top-k, dropping a sequence that reached the list of two
 * shards (the lower global index survives; seq_len = 0: no de-duplication).  world * k <= 8192. */
int64_t flexs_topk_select_workspace_bytes(void);
int flexs_topk_select_dev(const float *d_scores, int64_t n, int k, int64_t index_offset,
                          const uint8_t *d_rows, int row_len, int unique, float *d_top_scores,
                          int64_t *d_top_idx, uint8_t *d_top_rows, int *d_status, void *d_work,
                          void *stream);
int64_t flexs_screen_message_bytes(int k, int seq_len);
int flexs_screen_merge_dev(const void *d_gathered, int world, int k, int seq_len,
                           float *d_top_scores, int64_t *d_top_idx, uint8_t *d_top_rows,
                           void *stream);

/* ---- K3d: the screen's exchange step as one-sided stores over NVLink peer memory ----------
 * Alternative to the all_gather of screen.py for one-process-per-GPU jobs on one node (the
 * reference has no counterpart: adalead.py:171-175 ranks one in-process list).  Every rank owns
 * a mailbox  data [depth][world][msg_bytes] | flags [depth][world] uint32  allocated with
 * flexs_peer_alloc (cudaMalloc + a 64-byte cudaIpc handle the host side passes to the other
 * ranks, which map it with flexs_peer_open).  flexs_screen_push_dev writes this rank's message
 * into slot `slot` of EVERY mailbox (d_peer_bases: device array of `world` mailbox pointers, own
 * one included) and then stores the step number `seq` (> 0, increasing) into the slot's flag
 * with release semantics at system scope; it never waits.  flexs_screen_wait_dev blocks the
 * STREAM (not the host) until all `world` flags of `slot` in this rank's own mailbox have
 * reached `seq` (20 s watchdog: *d_status = 2 and a trapped launch instead of a hung GPU); the
 * slot's data block is then the rank-major layout flexs_screen_merge_dev takes.
 * msg_bytes: a multiple of 16 (flexs_screen_message_bytes is).                              */
int64_t flexs_peer_mailbox_bytes(int64_t msg_bytes, int world, int depth);
int flexs_peer_alloc(int64_t bytes, void **d_ptr, unsigned char *handle64);
int flexs_peer_open(const unsigned char *handle64, void **d_ptr);
int flexs_peer_close(void *d_ptr);
int flexs_peer_free(void *d_ptr);
int flexs_screen_push_dev(const void *d_msg, int64_t msg_bytes, int rank, int world, int slot,
                          int depth, uint32_t seq, const void *d_peer_bases, void *stream);
int flexs_screen_wait_dev(const void *d_mailbox, int64_t msg_bytes, int world, int slot,
                          int depth, uint32_t seq, int *d_status, void *stream);

/* ---- K5: candidate generation helpers -------------------------------------------------
 * flexs_mutate_dev replaces generate_random_mutant (sequence_utils.py:87-108) applied to
 * n parents at once: every residue is, with probability mu, replaced by a uniform draw
 * from the alphabet (which may re-draw the same residue, as the reference does).
 * RNG is Philox-4x32-10 keyed by (seed, flat position); the reference is unseeded, so
 * parity is distributional (tests check rates), not bitwise.                            */
int flexs_mutate_dev(const uint8_t *d_parents, int64_t n, int seq_len, int alphabet_size,
                     float mu, uint64_t seed, uint64_t subsequence, uint8_t *d_children,
                     void *stream);

/* flexs_argmax_decode_dev replaces CMAES._soln_to_string (cmaes.py:61-67) and the decode in
 * environments/dyna_ppo.py:144-147: d_x is [n, seq_len, row_stride] floats of which the
 * first `alphabet_size` of each row are compared; first maximum wins (np.argmax).         */
int flexs_argmax_decode_dev(const float *d_x, int64_t n, int seq_len, int row_stride,
                            int alphabet_size, uint8_t *d_idx, void *stream);

/* ---- K8: the density penalty of the DyNA-PPO environment ----------------------------------
 * Replaces DynaPPOEnvironment.sequence_density (flexs/environments/dyna_ppo.py:106-114) applied to a
 * whole batch: d_out[i] = sum over the n_seen rows o with 0 < d(new_i, o) <= radius of
 * d_seen_fitness[o] / d(new_i, o), d = Levenshtein distance (the reference calls editdistance.eval,
 * radius 2).  Rows are uint8[*, seq_len] residue indices; sums in float64, fixed order.            */
int flexs_edit_density_dev(const uint8_t *d_new, int64_t n_new, const uint8_t *d_seen,
                           const double *d_seen_fitness, int64_t n_seen, int seq_len, int radius,
                           double *d_out, void *stream);

/* ---- K4: training --------------------------------------------------------------------
 * Replaces keras Model.fit as called at keras_model.py:61-67 with the compile() of
 * cnn.py:56 / mlp.py:33: MSE loss, Adam(1e-3, 0.9, 0.999, 1e-7), batch_size, epochs,
 * shuffle every epoch, Dropout(0.25) active for the CNN (cnn.py:51).  Optimiser state
 * persists in the model across calls, as it does in the reference (explorer.py:157-160).
 * d_idx uint8[n, seq_len], d_labels float32[n].  h_losses (may be NULL) receives the
 * mean loss of each epoch.  All members are trained (ensemble.py:42-52).                 */
int flexs_model_fit_dev(flexs_model_t *m, const uint8_t *d_idx, const float *d_labels,
                        int64_t n, int batch_size, int epochs, uint64_t seed,
                        float *h_losses, void *stream);

/* One optimiser step on exactly the given batch with a caller-supplied dropout mask
 * (float32 [n, H] of 0/1, NULL = no dropout) — the unit the parity tests pin against
 * oracle.cnn_loss_and_grads + adam_update.  Returns the batch loss in *h_loss.           */
int flexs_model_train_step_dev(flexs_model_t *m, int member, const uint8_t *d_idx,
                               const float *d_labels, int64_t n, const float *d_dropout_mask,
                               float *h_loss, void *stream);

/* Adam moments (m, v; same arrays/layout as the weights; either pointer table may be NULL) and
 * the 1-based step count of one member — what keras `model.optimizer.get_weights()` exposes; the
 * parity tests read the gradient of a first step back as m / (1 - beta1).                       */
int flexs_model_get_optimizer_state(flexs_model_t *m, int member, float *const *h_m, float *const *h_v,
                                    int64_t *step);
/* Forget the optimiser state of every member (a freshly compiled Keras model).                  */
int flexs_model_reset_optimizer(flexs_model_t *m);

/* ---- K6: table landscapes ------------------------------------------
 * Ground-truth landscapes that are pure tables, evaluated on the device so a whole explorer
 * round can stay there.  d_seq is uint8[n, seq_len]: residue CHARACTERS when h_column_of_char
 * (256 host bytes: table column of each character, 0xFF = none) is non-NULL, else column
 * indices.  Outputs are float64, like the reference's.
 *
 * flexs_additive_score_dev replaces AdditiveAAVPackaging._fitness_function
 * (flexs/landscapes/additive_aav_packaging.py:101-118):
 *   out = max(0, (sum_i table[i, col(seq[i])] + offset) / denom + noise)
 * with offset = mfm * max_possible, denom = max_possible * (mfm + 1); characters without a
 * column contribute nothing (:104); the sum runs left to right in float64, bit-identical to the
 * reference's Python loop.  d_table float64[seq_len, ncols]; d_noise float64[n] or NULL.
 *
 * flexs_lookup_score_dev replaces TFBinding._fitness_function (flexs/landscapes/tf_binding.py:43-44):
 * the dict of all sequences is a dense float64 table of base**seq_len entries indexed by the
 * big-endian base-`base` number formed by the columns; a sequence with an unknown character is
 * reported as NaN (as is a key the caller left NaN in the table) — the host raises KeyError.  */
int flexs_additive_score_dev(const uint8_t *d_seq, int64_t n, int seq_len,
                             const uint8_t *h_column_of_char, int ncols, const double *d_table,
                             double offset, double denom, const double *d_noise, double *d_out,
                             void *stream);
int flexs_lookup_score_dev(const uint8_t *d_seq, int64_t n, int seq_len,
                           const uint8_t *h_column_of_char, int base, const double *d_table,
                           int64_t table_len, double *d_out, void *stream);

/* ---- K9: the VAE generator of CbAS / DbAS ---------------------------
 * Replaces flexs/utils/VAE_utils.py: VAEModel (:28-92: encoder Dense-ELU, Dropout(0.3), Dense-ELU,
 * BatchNorm, Dense-ELU, z_mean / z_log_var, sampling; decoder Dense-ELU x2, Dropout(0.3), Dense-ELU,
 * Dense-sigmoid; loss = sum_d BCE + KL), compile (:127: Adam 1e-4, clipvalue 0.5), fit (:141-151),
 * generate's decoder pass (:71-74) and calculate_log_probability (:189-217).  22 weight arrays in
 * Keras get_weights() order: W1 (L*A, I) b1 | W2 (I, I) b2 | BatchNorm gamma beta moving_mean
 * moving_variance | W3 b3 | Wm (I, Z) bm | Wv (I, Z) bv | W4 (Z, I) b4 | W5 b5 | W6 b6 | W7 (I, L*A) b7.
 * Sequences are residue indices uint8[n, seq_len]; the one-hot input is never built.
 *
 * flexs_vae_fit_dev: `epochs` passes over rows [0, n_train) of d_idx with sample weights d_weights
 * (the caller holds out the validation split, VAE_utils.py:148), mini-batches of batch_size reshuffled
 * every epoch, early stopping when the epoch's mean loss has not improved for `patience` epochs
 * (:139; 0 = never); h_losses[e] = mean loss of epoch e, *epochs_run = epochs actually run.  One
 * host synchronisation per epoch.  flexs_vae_train_step_dev is the parity unit: one optimiser step
 * on exactly the given batch with caller-supplied dropout masks (float [n, I], values 0 or 1/0.7)
 * and latent noise (float [n, Z]); flexs_vae_get_gradients reads the gradients it computed.
 * flexs_vae_decode_dev: decoder on d_z float[n, Z] -> d_out float[n, L*A] (inference mode).
 * flexs_vae_log_prob_dev: log-probability of reconstructing each sequence (:189-217, float64,
 * nan_to_num applied); d_eps float[n, Z] is the latent noise predict() draws (NULL: z = z_mean).  */
int flexs_vae_create(int device, int seq_len, int alphabet_size, int intermediate_dim, int latent_dim,
                     flexs_vae_t **out);
void flexs_vae_destroy(flexs_vae_t *v);
int flexs_vae_num_arrays(const flexs_vae_t *v);
int64_t flexs_vae_array_size(const flexs_vae_t *v, int i);
int flexs_vae_set_weights(flexs_vae_t *v, const float *const *h_arrays);
int flexs_vae_get_weights(flexs_vae_t *v, float *const *h_arrays);
int flexs_vae_get_gradients(flexs_vae_t *v, float *const *h_arrays);
int flexs_vae_reset_optimizer(flexs_vae_t *v);
int flexs_vae_train_step_dev(flexs_vae_t *v, const uint8_t *d_idx, const float *d_weights, int64_t n,
                             const float *d_mask1, const float *d_mask2, const float *d_eps,
                             float *h_loss, void *stream);
int flexs_vae_fit_dev(flexs_vae_t *v, const uint8_t *d_idx, const float *d_weights, int64_t n_train,
                      int batch_size, int epochs, int patience, uint64_t seed, float *h_losses,
                      int *epochs_run, void *stream);
int flexs_vae_decode_dev(flexs_vae_t *v, const float *d_z, int64_t n, float *d_out, void *stream);
int flexs_vae_log_prob_dev(flexs_vae_t *v, const uint8_t *d_idx, int64_t n, const float *d_eps,
                           double *d_logp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FLEXS_B200_H */
