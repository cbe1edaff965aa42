/*
 * focal_b200 -- C ABI of the B200-native FOCAL contrastive-loss hot path.
 *
 * The reference (tomoyoshki/focal) has no FFI: its hot path is the Python class
 * FOCALLoss (/root/reference/src/models/loss.py:8-218) called from
 * calc_contrastive_loss (/root/reference/src/train_utils/loss_calc_utils.py:9).  These entry
 * points are what a binding for that class would call; INTEGRATION.md shows the ctypes stub and the
 * drop-in `models/loss.py`.  Conventions: plain pointers and sizes only; every function returns 0 or a
 * negative FOCAL_E* code and never throws; nothing here allocates device memory or synchronises the
 * stream; all device pointers are caller-owned, 16-byte aligned, row-major contiguous fp32 unless
 * stated; all mutable state lives in the caller's workspace, so calls are re-entrant per workspace.
 */
#ifndef FOCAL_B200_H_
#define FOCAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOCAL_B200_ABI_VERSION 5
#define FOCAL_MAX_MODALITIES 8

enum {
  FOCAL_OK = 0,
  FOCAL_EINVAL = -1,      /* bad argument (null pointer, negative size, ...) */
  FOCAL_ESHAPE = -2,      /* B % S != 0, unsupported S / D / M (see focal_b200_strerror) */
  FOCAL_ECUDA = -3,       /* a CUDA runtime call failed (launch error, wrong arch, ...) */
  FOCAL_EWORKSPACE = -4   /* workspace too small / misaligned */
};

/* terms bit mask: which parts of loss.py:139-218 to evaluate */
enum { FOCAL_TERM_NCE = 1, FOCAL_TERM_ORTH = 2, FOCAL_TERM_TEMPORAL = 4, FOCAL_TERM_ALL = 7 };
/* precision of the Gram tiles (everything outside the tiles is fp32; accumulation is always fp32):
 *   FOCAL_PREC_BF16  bf16 operands, one tcgen05 kind::f16 pass: north_star's "bf16 mode" (gradients within 1e-2)
 *   FOCAL_PREC_FP32  north_star's "fp32 mode" (gradients within 2e-3): every operand element travels as a bf16 hi / lo
 *                    pair (16 significant bits) and every product is three kind::f16 passes hi*hi + hi*lo + lo*hi.
 *                    kind::tf32 cannot serve here: the B tile is the K-major operand of the Gram and the MN-major
 *                    operand of the gradient GEMM, and no shared-memory layout is legal for both views of 32-bit
 *                    elements (profiles/r2_tf32_probe.txt).  D <= 256. */
enum { FOCAL_PREC_BF16 = 0, FOCAL_PREC_FP32 = 1 };

/*
 * The values FOCALLoss reads from `args` (loss.py:11-23, 149, 163, 211-215) plus build-side options.
 *   B      rows of every feature tensor (= b sequences x S neighbouring windows; loss.py:152-155)
 *   S      dataset_config["seq_len"]           M  len(dataset_config["modality_names"])
 *   D      embedding width; shared = [0, D/2), private = [D/2, 2*(D/2))   (FOCALModules.py:51-57)
 *   seq_begin/seq_end   sequences whose rows this call owns (multi-GPU row shard); [0, B/S) on one GPU.
 *          Operands are always built for all B rows; losses and gradients only for the owned rows.
 */
typedef struct FocalCfg {
  int32_t B, S, M, D;
  float temperature;     /* FOCAL.temperature (already resolved for args.model) */
  float margin;          /* FOCAL.inter_rank_margin */
  float w_shared, w_private, w_orth, w_rank; /* the four loss weights, loss.py:211-215 */
  int32_t no_private;    /* args.tag == "noPrivate": shared InfoNCE on full-width features (loss.py:163-170) */
  int32_t need_grad;     /* 0: forward only (eval_functions.py:72-80 runs under no_grad) */
  int32_t terms;         /* FOCAL_TERM_* mask; FOCAL_TERM_ALL for the reference loss */
  int32_t precision;     /* FOCAL_PREC_* */
  int32_t seq_begin, seq_end;
  int32_t num_sms;       /* 0 = query the current device */
  /* Row-blocked inputs (what an all-gather of per-rank [2M][B/R][D] buffers produces): row i of feature tensor t
   * lives at feats[t] + (i / in_block_rows) * in_block_stride + (i % in_block_rows) * D (in floats).
   * in_block_rows == 0 means plain contiguous [B, D] tensors.  Sequences never straddle blocks. */
  int32_t in_block_rows;
  int32_t in_block_stride;
  /* local_rows != 0 (only with focal_b200_loss_sharded): feats[t] / grads[t] hold just the owned rows
   * [seq_begin*S, seq_end*S) of the B-row tensors, i.e. what a rank of a row-sharded job has in hand. */
  int32_t local_rows;
  /* indirect_ptrs != 0: the feats / grads / loss5 arguments of the stage functions are ignored; the kernels read them
   * from a table in the workspace that focal_b200_set_ptrs fills.  This is what makes the launch sequence replayable
   * from a CUDA graph with different (freshly allocated) inputs and outputs every step. */
  int32_t indirect_ptrs;
} FocalCfg;

/* Byte offsets (from the workspace base) and extents of the buffers a host may need to look at:
 * the InfoNCE row sums exchanged between ranks, and intermediates the parity tests compare. */
typedef struct FocalWsInfo {
  size_t total_bytes;
  int32_t b, bpad, Bpad, n_problems, n_ops, kb_full;
  size_t rowsum_off;     /* fp32 [n_problems][S][2][bpad]  sum_{j != k} exp(s_kj), position-major          */
  size_t rowsum_bytes;
  size_t cnt_off;        /* int32 [2M][bpad]: active hinge count per sequence (temporal); one copy per      */
  size_t cnt_bytes;      /* stream-K piece: copy k at cnt_off + k * cnt_piece_stride, k < n_pieces_tmp; sum them */
  size_t mintra_off;     /* fp32 [2M][Bpad]: m_II of the row's sequence                                    */
  size_t lossparts_off;  /* double [8]: total, shared, private, orth, temporal (un-weighted sums) ...      */
  size_t dz_off, dz_bytes;     /* fp32 InfoNCE operand-gradient accumulators                               */
  size_t dx_off, dx_bytes;     /* fp32 [2M][Bpad][kb_full*64] temporal gradient accumulators               */
  size_t cnt_piece_stride;     /* bytes between the per-piece copies of cnt                                */
  int32_t n_pieces_nce, n_pieces_tmp; /* most pieces a 128-row block is cut into (stream-K over column tiles) */
} FocalWsInfo;

int focal_b200_abi_version(void);
const char* focal_b200_strerror(int code);

/* Validates cfg and reports workspace size + layout.  Replaces nothing in the reference (which allocates
 * its [S,N,N,d] temporaries per call, loss.py:74); the caller allocates once and re-uses. */
int focal_b200_workspace_info(const FocalCfg* cfg, FocalWsInfo* info);

/* Stage 1 -- row prologue over ALL rows: L2 norms (eps 1e-8, loss.py:15/74), normalised + pre-scaled bf16
 * InfoNCE operands in position-major swizzled tiles (loss.py:66-73), bf16 temporal operands + squared norms
 * (loss.py:117), exact intra-sequence mean distances m_II (loss.py:118-124), orthogonality terms
 * (loss.py:89-106) and positive-pair logits (loss.py:75-79) of the owned rows.
 * feats: 2M device pointers, order view-major then modality: [v0m0, v0m1, ..., v1m0, ...], each fp32 [B, D]. */
int focal_b200_prologue(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, void* stream);

/* Stage 2 -- InfoNCE row sums sum_{j != k} exp(s_kj) of the owned rows (loss.py:74-85), tcgen05 Gram + exp2. */
int focal_b200_nce_rowsum(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream);

/* Stage 2b -- reduce column-split partial sums, take logs, publish 1/rowsum.  When `all_rows` is non-zero the
 * reciprocal table is rebuilt for every row (after a multi-GPU exchange filled in the other ranks' row sums). */
int focal_b200_nce_lse(const FocalCfg* cfg, void* ws, size_t ws_bytes, int all_rows, void* stream);

/* Stage 3 -- InfoNCE backward of the owned rows (autograd of loss.py:74-85): recompute logit tiles,
 * W = E (1/r_k + 1/r_j), second UMMA W @ Z_J into TMEM. */
int focal_b200_nce_grad(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream);

/* Stage 4 -- temporal ranking term (loss.py:108-137), forward and backward fused in one pass over the
 * B x B distance tiles of the owned rows. */
int focal_b200_temporal(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream);

/* Stage 5 -- assemble: loss5 = {total, shared, private, orth, temporal} (device, fp32[5], un-weighted parts,
 * weighted total; partial over the owned rows) and, when cfg->need_grad, d total / d feats into grads
 * (2M device pointers, fp32 [B, D]; only the owned rows are written). */
int focal_b200_finalize(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, float* loss5,
                        float* const* grads, void* stream);

/* Indirect mode (cfg->indirect_ptrs): store this step's caller pointers in the workspace table (one tiny launch on
 * `stream`).  grads may be NULL when !cfg->need_grad.  Call it before every replay of a captured launch sequence. */
int focal_b200_set_ptrs(const FocalCfg* cfg, void* ws, size_t ws_bytes, const float* const* feats, float* loss5,
                        float* const* grads, void* stream);

/* All stages in order on one stream (single-GPU call of FOCALLoss.forward + backward). */
int focal_b200_loss(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, float* loss5,
                    float* const* grads, void* stream);

/*
 * Row-sharded multi-GPU path over NVLink peer memory (one process per GPU, all on one NVSwitch box).  The reference is
 * single-GPU; semantics: R ranks each hold B/R rows (whole sequences, rank-major); every rank gets the
 * loss of the global batch and d loss_global / d (its own rows).  No collective library on the data path: each rank
 * allocates its workspace with focal_b200_peer_alloc, ships the 64-byte handle to the other ranks (any host channel),
 * maps theirs with focal_b200_peer_open, and the kernels exchange what they produce with plain stores into the
 * peers' workspaces: the prologue runs on the owned rows only and writes their bf16 operands into every workspace;
 * nce_lse publishes the owned rows' row sums the same way; the loss partials are all-reduced inside the last kernel.
 * Three device-side barriers (flags in the workspaces, bounded spin, trap on timeout) order the phases.
 */
#define FOCAL_MAX_PEERS 8
typedef struct FocalPeers {
  int32_t rank, world;
  void* ws[FOCAL_MAX_PEERS];   /* this process's mapping of every rank's workspace; ws[rank] is its own */
  void* mc;                    /* optional (NULL = none): an NVSwitch MULTICAST mapping of the same workspaces -- a store to
                                * mc + off lands at ws[r] + off of every rank r.  With it the kernels send the operands and
                                * row sums of the owned rows once (multimem.st) instead of once per peer.  The workspaces
                                * then come from a multicast-capable allocator (e.g. torch symmetric memory), zeroed, not
                                * from focal_b200_peer_alloc. */
} FocalPeers;

/* cudaMalloc + zero-fill + cudaIpcGetMemHandle; the workspace must come from here so that peers can map it. */
int focal_b200_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int focal_b200_peer_open(const unsigned char handle[64], void** ptr);
int focal_b200_peer_close(void* ptr);
int focal_b200_peer_free(void* ptr);

/* All stages of one rank's share (cfg->seq_begin/seq_end = owned sequences, cfg->local_rows = 1, cfg identical on all
 * ranks otherwise).  Every rank must call it the same number of times, in step; ws_bytes as for focal_b200_loss.
 * loss5 receives the GLOBAL loss.  Returns FOCAL_ESHAPE for shapes off the vectorised row-kernel path (D/2 a multiple
 * of 32, S in {1, 2, 4}, no noPrivate) -- callers then use the staged functions with a collective library. */
int focal_b200_loss_sharded(const FocalCfg* cfg, const float* const* feats, const FocalPeers* peers, size_t ws_bytes,
                            float* loss5, float* const* grads, void* stream);
/* Diagnostics of the row-sharded path (no reference counterpart): with FOCAL_B200_STAGE_TIMES=1 in the environment,
 * eager (not stream-captured) calls of focal_b200_loss_sharded record CUDA events between their launches; this returns
 * the milliseconds of each stage of the last call (prologue, nce_rowsum, nce_lse, temporal, nce_grad, finalize --
 * including what a stage waits for from the peers).  Synchronises.  Returns the number of values written (0: disabled). */
int focal_b200_debug_stage_times(float* out_ms, int n);

/*
 * Frequency-domain input stage: the part of Augmenter.fft_preprocess after torch.fft.fft
 * (/root/reference/src/data_augmenter/Augmenter.py:141-158: view_as_real + permute + reshape to [b, 2c, i, s]) fused
 * with PhaseShiftAugmenter's rotation (/root/reference/src/data_augmenter/PhaseShiftAugmenter.py:35-57).
 *   in   interleaved != 0: complex64 [n_bc][plane] (re, im) pairs, e.g. the cuFFT output viewed as float;
 *        interleaved == 0: planar fp32 [n_bc][2][plane] (what PhaseShiftAugmenter receives)
 *   out  planar fp32 [n_bc][2][plane]:  re' = re cos - im sin,  im' = re sin + im cos
 * n_bc = batch x complex channels, plane = intervals x spectrum length (a multiple of 4); 16-byte aligned pointers.
 */
int focal_b200_spectrum_rotate(const float* in, float* out, long long n_bc, int plane, int interleaved, float cos_angle,
                               float sin_angle, void* stream);

/*
 * k-nearest-neighbour evaluation: replaces sklearn's KNeighborsClassifier().fit / .predict in
 * /root/reference/src/train_utils/knn.py:22-42 and eval_functions.py:65-97 (Euclidean, uniform weights, majority vote,
 * ties to the smallest label; k <= 16, labels in [0, n_classes), n_classes <= 32) without leaving the device.
 *   train fp32 [n_train, dim], labels int32 [n_train], query fp32 [n_query, dim]
 *   dist_ws fp32 [n_query, n_train] scratch (squared distances are left there), out int32 [n_query] predicted labels,
 *   neighbours int32 [n_query, k] indices of the k nearest training rows in order, or NULL
 */
int focal_b200_knn_predict(const float* train, const int32_t* labels, int n_train, const float* query, int n_query,
                           int dim, int k, int n_classes, float* dist_ws, int32_t* out, int32_t* neighbours,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FOCAL_B200_H_ */
