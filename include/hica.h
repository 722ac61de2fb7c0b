/*
 * hica.h - C ABI of libhica.so, the B200-native (sm_100a) replacement for the
 * contact-map scoring hot path of AO33/HiC_Genome_Assembler.
 *
 * The reference has no FFI or plugin API of its own (it is 100 % Python and
 * calls numpy/scipy/numba in-process).  The boundary is therefore the set of
 * module-level functions its two pipelines call on the hot path
 * (SURVEY.md section 8b); every entry point below names the reference
 * function/lines it replaces (paths relative to /root/reference/HIC_ASSEMBLER/).
 * INTEGRATION.md shows the ctypes stub that rebinds those functions.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is
 *     available from hica_last_error(h) (or hica_last_error(NULL) for
 *     hica_create failures);
 *   - the caller owns all host buffers (C-contiguous, must outlive the call);
 *     the library owns all device memory behind the opaque handle;
 *   - calls are synchronous with respect to the host: when a function returns
 *     its host outputs are valid;
 *   - one handle per GPU; a handle is not re-entrant;
 *   - there is NO CPU fallback: without a CUDA device hica_create fails.
 */
#ifndef HICA_H
#define HICA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hica_ctx *hica_handle;

#define HICA_OK 0
#define HICA_ERR_CUDA 1
#define HICA_ERR_ARG 2
#define HICA_ERR_STATE 3
#define HICA_ERR_NOMEM 4

/* ---- lifetime ----------------------------------------------------------- */
int hica_version(void);
int hica_create(int device, hica_handle *out);
int hica_destroy(hica_handle h);
const char *hica_last_error(hica_handle h);
/* number of kernel launches issued through this handle so far */
int64_t hica_launch_count(hica_handle h);
/* Wall-clock independent stage timers: CUDA-event milliseconds of the last
 * call of each stage (index = HICA_T_*), measured on the handle's stream. */
enum {
    HICA_T_UPLOAD = 0, HICA_T_DISTANCE, HICA_T_LINKAGE, HICA_T_LEAFORDER, HICA_T_RANK,
    HICA_T_BREAKPOINTS, HICA_T_FILTER, HICA_T_P2_SCORE, HICA_T_RANK_PRESORT, HICA_T_COUNT
};
int hica_get_timers(hica_handle h, double *ms /* HICA_T_COUNT */);
/* counters of the last hica_p1_linkage: nearest-neighbour searches, SM cycles spent in searches, in merges,
 * the search cycles split into row load / in-CTA reduction / cross-CTA exchange, searches answered from the
 * nearest-neighbour cache, reserved (8 values; "searches" counts the ones done in full) */
int hica_linkage_stats(hica_handle h, int64_t *out8);
/* host-side phase times (ms) of the last hica_p2_order_groups call: sizing, arena cudaMalloc, setup, enqueue +
 * round synchronisation, read-back, cudaFree (16 slots, the rest reserved) */
int hica_debug_stats(hica_handle h, double *out16);

/* ---- Part 1: ICE matrix -> leaf order -> cut indices ---------------------- */

/* Copy the dense n x n ICE matrix (after removeRows, scaffoldToChromosomes.py
 * :1116-1118) to the device.  A_host may be pageable or pinned. */
int hica_p1_upload(hica_handle h, const double *A_host, int64_t n);

/* Same for a SYMMETRIC matrix (the contact map the reference builds is: it writes adjMat[i][j] and adjMat[j][i]
 * together, scaffoldToChromosomes.py:1069-1082): only the upper triangle (row blocks right of the diagonal,
 * ~53 % of the bytes) crosses PCIe, the lower triangle is mirrored on the device.  Elements below the diagonal
 * of A_host are never read. */
int hica_p1_upload_symmetric(hica_handle h, const double *A_host, int64_t n);

/* Same, but the matrix is already in device memory (row-major, ld = n). The
 * library copies it; the caller keeps ownership of A_dev. */
int hica_p1_upload_device(hica_handle h, const double *A_dev, int64_t n);

/* ---- row-block sharding over the GPUs of one box -------------
 * One process and one handle per GPU.  The contact matrix that removeRows
 * (scaffoldToChromosomes.py:1116-1118) leaves in host memory is cut into
 * contiguous row blocks; every GPU uploads its own block (so the N PCIe links
 * work in parallel), the caller assembles the full matrix on every GPU with a
 * collective over NVLink on the buffer hica_p1_device_buffer(h, 0) returns
 * (torch.distributed / NCCL all-gather or broadcasts, in place), and - for a
 * symmetric upload - hica_p1_mirror_lower fills the lower triangle.
 *   hica_p1_alloc        allocate the n x n matrix without filling it
 *   hica_p1_upload_rows  rows [r0, r1): rows_host points at row r0 (ld = n);
 *                        symmetric != 0: only columns >= the row's 512-row
 *                        sub-block start cross PCIe (the matrix must be
 *                        symmetric; finish with hica_p1_mirror_lower)
 *   hica_p1_device_buffer  which = 0: the matrix (fp64, n*n); which = 1: the
 *                        rank matrix (uint16 for n <= 65535, else int32, n*n).
 *                        The pointer stays valid until the next upload/alloc of
 *                        a different n (matrix) or a larger rank matrix. */
int hica_p1_alloc(hica_handle h, int64_t n);
int hica_p1_upload_rows(hica_handle h, const double *rows_host, int64_t r0, int64_t r1, int symmetric);
int hica_p1_mirror_lower(hica_handle h);
int hica_p1_device_buffer(hica_handle h, int which, void **ptr, int64_t *bytes, int32_t *elem_bytes);

/* binList[i].rowSum (removeRows, :134: builtin left-to-right sum).  Pass the
 * host values the reference computed, or NULL to compute them on the device in
 * the same left-to-right order. */
int hica_p1_set_row_sums(hica_handle h, const double *row_sum_host_or_null);

/* convertMatrix(distance=True) :138-147 + squareform :194: numpy-pairwise row
 * sums, then the mirrored working matrix W[i][j] = W[j][i] = (1 - A_ij/s_i) + 1
 * (i < j) that the linkage consumes. */
int hica_p1_distance(hica_handle h);

/* The full asymmetric D (what convertMatrix(distance=True) returns), n x n. */
int hica_p1_get_distance(hica_handle h, double *D_host);
/* The numpy-pairwise row sums s_i used by the distance transform. */
int hica_p1_get_pairwise_row_sums(hica_handle h, double *s_host);
int hica_p1_get_row_sums(hica_handle h, double *row_sum_host);

/* Replace the working matrix by the strict upper triangle of a caller-supplied
 * square matrix D (the seam averageClusterNodes(adjacencyMatrix, ...) :187 when
 * it is handed a matrix that did not come from hica_p1_distance). */
int hica_p1_upload_distance(hica_handle h, const double *D_host, int64_t n);

/* averageClusterNodes :187-208 = hierarchy.average (nn_chain, stable sort by
 * height, union-find relabel) + dendrogram(count_sort='ascending') leaf order.
 * Z_host: (n-1) x 4 doubles or NULL; leaves_host: n int32 or NULL.  The leaf
 * order also becomes the handle's current row/column order. */
int hica_p1_linkage(hica_handle h, double *Z_host, int32_t *leaves_host);

/* Set the row/column order explicitly (readDengrogramLeavesFromFile :222,
 * reorderMatrix :157-163 by indirection). */
int hica_p1_set_order(hica_handle h, const int32_t *order);

/* Materialise reordered matrices for the Python seams:
 *   what = 0: D[:, o][o]              (reorderMatrix :161)
 *   what = 1: S = rowSum*(1-(D-1))    (convertMatrix(similarity=True) :148-149) */
int hica_p1_get_reordered(hica_handle h, int what, double *out_host);

/* numpy.argsort(S, axis=1)[:, ::-1] :1132 on the device; ties resolve as a
 * stable ascending sort reversed (descending column index). */
int hica_p1_rank(hica_handle h);
/* Start the value sort of :1132 BEFORE the leaf order is known (the values of a
 * row do not depend on the order, only their column labels do): every row is
 * sorted in matrix order on a second, low-priority stream and the call returns
 * at once.  hica_p1_linkage starts it by itself next to the linkage chain
 * (which keeps 16 of the 148 SMs busy for its whole duration); call it
 * explicitly on GPUs that wait for another GPU's linkage.  A later hica_p1_rank
 * / hica_p1_rank_rows then only relabels the columns and orders exact ties
 * (HICA_T_RANK = that part, HICA_T_RANK_PRESORT = the overlapped sort).
 * Needs the ICE matrix, hica_p1_set_row_sums and hica_p1_distance; n <= 65535
 * (16-bit column labels; rows of up to 16,384 entries are sorted by the radix
 * kernels, longer ones chunk-wise by the bitonic network).
 * HICA_RANK_OVERLAP=0 disables it, HICA_RANK_OVERLAP_CHUNKED=0 only for the
 * chunk-wise path. */
int hica_p1_rank_presort(hica_handle h);

/* Row-block sharded form of hica_p1_rank (:1132): rank rows
 * [r0, r1) of the reordered matrix only.  The other row blocks of the rank
 * matrix are written by the caller's collective on hica_p1_device_buffer(h, 1);
 * hica_p1_rank_commit then declares the rank matrix complete. */
int hica_p1_rank_rows(hica_handle h, int64_t r0, int64_t r1);
int hica_p1_rank_commit(hica_handle h);
/* Seam input for :1132 when the similarity matrix S was computed elsewhere: upload it as is (identity
 * order); hica_p1_rank then sorts its rows directly. */
int hica_p1_upload_similarity(hica_handle h, const double *S_host, int64_t n);
int hica_p1_get_rank(hica_handle h, int64_t *R_host);
/* Seam input: an argsorted matrix computed elsewhere (n x n int64). */
int hica_p1_set_rank(hica_handle h, const int64_t *R_host, int64_t n);

/* One evaluation of the per-row significance vector.
 *   mode 0: find_matrix_pvalue_breakpoints :446-469 - rows i in (start, n),
 *           c = i-start, x = #{k<c : start <= R[i][k] <= i},
 *           sig = !(sf(x-1; M, c, c) >= psig)            (NaN -> 1)
 *           out arrays have n-start entries, entry 0 = 0.
 *   mode 1: filter_noisy_breakpoints :617-636 with cut index `cut` - rows
 *           ii in [start, n), x = #{k<cut-start : start <= R[ii][k] <= cut},
 *           sig = sf(x-1; M, local, local) < psig        (NaN -> 0), rows
 *           beyond start + int(n/5) are 0.
 * sig_host / x_host may be NULL.  *ambiguous receives the number of decisions
 * whose p-value lies within the library's safety margin of psig. */
int hica_p1_window_sig(hica_handle h, int64_t start, int64_t cut, int64_t M, double psig, int mode,
                       uint8_t *sig_host, int64_t *x_host, int64_t *ambiguous);

/* pre_process_all_matrix_breakpoints :513-551 (first-stage psig is the literal
 * .05 as in the reference, :535).  *ambiguous = the number of significance
 * decisions that could come out differently in the reference: p-values within
 * 1e-7 of the significance level, plus decisions that depend on the order
 * numpy.argsort gives EXACTLY TIED similarities (:1132, implementation-defined;
 * see hica_p1_tie_report). */
int hica_p1_breakpoints(hica_handle h, int64_t min_size, double min_frac,
                        int64_t *cuts_host, int64_t capacity, int64_t *n_cuts, int64_t *ambiguous);

/* filter_noisy_breakpoints :553-727. */
int hica_p1_filter(hica_handle h, const int64_t *cuts_in, int64_t n_in, double psig,
                   int64_t *cuts_out, int64_t *n_out, int64_t *ambiguous);

/* Tie-order sensitivity of the last hica_p1_breakpoints / hica_p1_filter calls.
 * The rank matrix replaces numpy.argsort(S, axis=1)[:, ::-1] (:1132), whose order
 * of exactly tied values is implementation-defined (this library: stable
 * ascending, reversed).  A windowed count x = #{k < c : lo <= R[i][k] <= hi}
 * (:455-459, :631) depends on that order only when position c falls inside a run
 * of tied values; the cut kernels then bound the count over all orders of the
 * run, x_lo <= x <= x_hi, and count the decisions that differ between the two.
 *   out4[0]  1 when tie information exists for the current rank matrix (it does
 *            not for a rank matrix supplied through hica_p1_set_rank)
 *   out4[1]  tie-sensitive decisions in the last hica_p1_breakpoints (-1: n/a)
 *   out4[2]  ... in the last hica_p1_filter
 *   out4[3]  rows of the rank matrix that contain a tie (-1: not counted) */
int hica_p1_tie_report(hica_handle h, int64_t *out4);

/* P(X >= x) decisions `sf(x-1; M, n, N) < psig` for a batch of scalar queries
 * (hyper_geom :352-368); out[i] = 1/0, or 2 when inside the safety margin. */
int hica_hypergeom_less(hica_handle h, const int64_t *x, const int64_t *M, const int64_t *n,
                        const int64_t *N, int64_t count, double psig, uint8_t *out);

/* ---- Part 2: ordering score over candidate batches ------------------------ */

/* The genome (or chromosome) contact matrix of orderGenome :607 (raw ICE
 * values).  n x n doubles. */
int hica_p2_upload(hica_handle h, const double *G_host, int64_t n);
/* Use the Part1 matrix already resident in this handle. */
int hica_p2_use_p1_matrix(hica_handle h);

/* giveNewAdjMat :297-308: select the working sub-matrix S = G[ix_(idx, idx)]
 * (device-side gather).  Subsequent segment/bin indices refer to S. */
int hica_p2_select(hica_handle h, const int32_t *idx, int32_t L);

/* Scaffolds as segments of S: seg_ptr has nseg+1 entries into seg_bins, which
 * lists each scaffold's bin indices in "+" (ascending bin id) reading order. */
int hica_p2_set_segments(hica_handle h, const int32_t *seg_ptr, const int32_t *seg_bins, int32_t nseg);

/* total = sum_k numpy.trace(M, offset=k) (orderGenome.py:343, 448, 506) for the
 * matrix laid out by `cand` (nslots entries, each (segment << 1) | flip). */
int hica_p2_total(hica_handle h, const int32_t *cand, int32_t nslots, double *total);

/* costFunction_numba :184-191 over a batch of explicit candidates: cands is
 * B x nslots, each entry (segment << 1) | flip.  Returns the first strict
 * maximum above `carried` (best_idx = -1 when none), exactly the accept rule of
 * :349, :464, :535.  scores_or_null receives all B scores when non-NULL. */
int hica_p2_score_candidates(hica_handle h, const int32_t *cands, int64_t B, int32_t nslots,
                             double total, double carried, double *best, int64_t *best_idx,
                             double *scores_or_null);

/* Window enumeration on the device (bruteForceBestScore :432-473 and one window
 * of scanOrdering :513-541): the candidate with index c in [cand_lo, cand_hi)
 * is base[0:win_start] + arrangement(c) + base[win_start+win_n:], where
 * arrangement(c) places window scaffold order_table[c / n_orients][j] with
 * orientation orient_table[c % n_orients][j] (0 = '+', 1 = '-') at window slot
 * j.  base is nslots entries of (segment << 1) | flip.  best_idx is absolute. */
int hica_p2_score_window(hica_handle h, const int32_t *base, int32_t nslots, int32_t win_start, int32_t win_n,
                         const uint8_t *order_table, int32_t n_orders, const uint8_t *orient_table,
                         int32_t n_orients, int64_t cand_lo, int64_t cand_hi, double total, double carried,
                         double *best, int64_t *best_idx, double *scores_or_null);

/* orderGenome :591-628 / orderChromosome :551-586 for a batch of independent chromosome groups, entirely
 * on the device (one CUDA stream per group; brute force, greedy insertion and the sliding-window scan are
 * stream-ordered kernels whose accept/apply decisions are taken on the device).
 *   grp_ptr   ngroups+1 offsets into the scaffold arrays; the scaffolds of a group are listed in the
 *             reference's pool order (bin count descending, stable - initiateBinsAndScaffolds :279)
 *   seg_ptr   total_scaffolds+1 offsets into bin_rows
 *   bin_rows  genome-matrix row of every bin, each scaffold in '+' (ascending bin id) reading order
 *   ord_*_tabs / ori_*_tabs: the reference's enumeration tables (orders after removeReverseDuplicates
 *             :396-411, orientations of plusMinusPerms :413-430, 0 = '+', 1 = '-'); the brute-force tables
 *             for k = 1..8 scaffolds are concatenated, table k occupying [off[k], off[k+1]) bytes
 *             (rows of k entries; empty when unused), the scan tables have `scan` entries per row.
 *   out_layout  total_scaffolds entries: per group the final order as (local scaffold index << 1) | flip
 *   out_stats   ngroups x 4: closed-form candidate evaluations, exact re-scorings, 1 if the group's matrix
 *               needed full exact scoring (negative/asymmetric entries), 1 if no ordering scored above 0 */
int hica_p2_order_groups(hica_handle h, int32_t ngroups, const int32_t *grp_ptr, const int32_t *seg_ptr,
                         const int32_t *bin_rows, int32_t n_scaffolds, int32_t scan,
                         const uint8_t *ord_brute_tabs, const int32_t *ord_brute_off,
                         const uint8_t *ori_brute_tabs, const int32_t *ori_brute_off,
                         const uint8_t *ord_scan, int32_t n_ord_scan, const uint8_t *ori_scan, int32_t n_ori_scan,
                         int32_t *out_layout, double *out_cost, int32_t *out_rounds, int64_t *out_stats);

/* Explicit bin orders (B x L int32 into S): the most general form, used by
 * tests and by the costFunction_numba seam. */
int hica_p2_score_orders(hica_handle h, const int32_t *orders, int64_t B, int32_t L, double total,
                         double carried, double *best, int64_t *best_idx, double *scores_or_null);

/* ---- 10 kb configuration (BASELINE.json configs[4]): fp32 storage, row-block sharding over the GPUs ----------
 * 160,000 bins: the ICE matrix is 102 GB in fp32 and is cut into contiguous row blocks, one per GPU (one process and
 * one handle per GPU); the sequential linkage runs on ONE GPU, which additionally holds the mirrored working matrix
 * (102 GB).  Replaces convertMatrix(distance=True) :138-147 + squareform :194 + hierarchy.average :197 for a matrix
 * the reference itself cannot hold.  Arithmetic is fp64 on fp32 storage (one rounding per stored value); parity is
 * against this library's own fp64 path at <= 32,000 bins (stated tolerance, tests/test_gpu_shard32.py).
 *   hica_s32_alloc      rows [r0, r1) of the n x n matrix on this GPU (+ the working matrix when
 *                       with_linkage_matrix != 0)
 *   hica_s32_buffer     device pointer / bytes of which = 0: the row block (fp32, (r1-r0) x n, filled by the caller
 *                       - e.g. as a zero-copy torch tensor), 1: the working matrix (fp32, n x ld, ld = n rounded up
 *                       to 16), 2: the n row sums (fp64; the caller all-gathers the blocks the other GPUs computed)
 *   hica_s32_row_sums   s_i of the own rows (fp64 accumulation)
 *   hica_s32_build_w    rows [r0, r1) of the mirrored working matrix, W[i][j] = (1 - A_ij / s_min(i,j)) + 1, written
 *                       to W_dst (row-major, leading dimension ld; matrix row i goes to row i - dst_row0): NULL = this
 *                       GPU's own working matrix, otherwise e.g. the linkage GPU's (dst_row0 = 0), mapped with
 *                       hica_s32_ipc_open - the transform and the gather over NVLink are then ONE kernel -
 *                       or a staging buffer of r1 - r0 rows (dst_row0 = r0) for a transport without peer mappings
 *   hica_s32_ipc_*      CUDA IPC export / open / close of a buffer (64-byte handle), for one process per GPU
 *   hica_s32_linkage    hierarchy.average + dendrogram leaf order on the fp32 working matrix (Z may be NULL) */
int hica_s32_alloc(hica_handle h, int64_t n, int64_t r0, int64_t r1, int with_linkage_matrix);
int hica_s32_buffer(hica_handle h, int which, void **ptr, int64_t *bytes);
int hica_s32_row_sums(hica_handle h);
int hica_s32_build_w(hica_handle h, void *W_dst, int64_t ld, int64_t dst_row0);
int hica_s32_ipc_export(hica_handle h, int which, void *handle64);
int hica_s32_ipc_open(hica_handle h, const void *handle64, void **ptr);
int hica_s32_ipc_close(hica_handle h, void *ptr);
int hica_s32_linkage(hica_handle h, double *Z_host, int32_t *leaves_host);

/* ---- loaders either side of the hot path ------------------ */

/* buildAdjacencyMatrix (scaffoldToChromosomes.py:70-98): the sparse HiC-Pro
 * matrix text ("id1<TAB>id2<TAB>value" per line) -> the dense symmetric n x n
 * fp64 matrix, in host memory (no GPU involved; the file is mmap-ed and parsed
 * by `threads` threads, 0 = all cores).
 *   index_of_id  max_id+1 entries: matrix row of HiC-Pro bin id, or -1 when the
 *                bin was dropped by initiateLoci (:35-68) - such lines are
 *                skipped as in :86-89
 *   A_host       n*n doubles owned by the caller (e.g. the pinned buffer
 *                hica_p1_upload* reads from); zero-filled, then [i][j] and
 *                [j][i] are set for every line, the last line winning (:90-91)
 *   n_lines / n_edges  lines parsed / lines applied ("Edges added", :97)
 * A malformed line fails the call (the reference raises); the message, with
 * the 1-based line number, is returned by hica_io_last_error(). */
int hica_io_read_matrix(const char *path, const int32_t *index_of_id, int64_t max_id, int64_t n,
                        double *A_host, int threads, int64_t *n_lines, int64_t *n_edges);
const char *hica_io_last_error(void);

/* initiateLoci (scaffoldToChromosomes.py:35-68, orderGenome.py:30-63): the
 * HiC-Pro bed file (chrom, start, stop, bin id) read in step with the bias file
 * -> the bins that survive (bias != "nan", id in the optional set), as columns.
 * Host side, no GPU involved.  Two steps: hica_io_read_loci parses and keeps the
 * result in the calling thread; hica_io_loci_fetch copies it into caller
 * buffers of n_bins (chrom_off: n_bins + 1, chrom_buf: chrom_bytes) entries and
 * releases it.  chrom_off[i]..chrom_off[i+1] delimit the chromosome name of bin
 * i inside chrom_buf.
 *   keep_ids_sorted / n_keep  binID_dict (:53-55) as an ascending id list;
 *                             n_keep < 0: no filter
 * Universal newlines, int() / float() parsing and the error cases of the
 * reference (a malformed bed line raises there: the call fails here, message
 * with the line number in hica_io_last_error()) are described in hica_io.cu. */
int hica_io_read_loci(const char *bed_path, const char *bias_path, const int64_t *keep_ids_sorted,
                      int64_t n_keep, int64_t *n_bins, int64_t *chrom_bytes);
int hica_io_loci_fetch(int64_t *ids, int64_t *starts, int64_t *stops, double *bias, int64_t *chrom_off,
                       char *chrom_buf);

/* logTransformMatrix (scaffoldToChromosomes.py:165-183 with offset = 1,
 * orderGenome.py:160-178 with offset = 0): out[i] = 0 where in[i] == 0, else
 * log(in[i] + offset) / log(base) (reverse = 0) or pow(base, in[i]) - offset
 * (reverse != 0), with the host libm the reference's math.log / ** use;
 * `count` entries, in == out allowed.  Fails ("math domain error") where the
 * reference's math.log would raise. */
int hica_io_log_transform(const double *in, double *out, int64_t count, double base, double offset,
                          int reverse, int threads);

/* The colour-scale clip of plotContactMap: numpy.percentile(adjMat, lP) and
 * numpy.percentile(adjMat, hP) (plotContactMaps.py:30-31) over all n*n values.
 * A percentile does not depend on the order of the elements, so the resident
 * matrix answers it without a read-back: exact order statistics by a radix
 * select over the order-preserving 64-bit keys (6 passes over the matrix, any
 * number of ranks, four per sweep).
 *   what  0: D = convertMatrix(distance=True) of the resident matrix (:147)
 *         1: S = convertMatrix(similarity=True) (:149), row sums as set
 *         2: the resident matrix as stored
 *   k     nk zero-based ranks in ascending order of value, among the elements
 *         that are not NaN; out[i] = that element, bit for bit (-0.0 counts
 *         as +0.0); a rank outside [0, n*n - nan_count) gives NaN
 *   nan_count  elements that are NaN (numpy.percentile then returns nan)
 * The caller interpolates as numpy does (two neighbouring ranks per
 * percentile; hic_genome_assembler_b200/plot.py restates numpy's arithmetic). */
int hica_p1_order_stats(hica_handle h, int what, int32_t nk, const int64_t *k, double *out,
                        int64_t *nan_count);
/* The same for `count` doubles in host memory (e.g. the closing
 * convertMatrix(logTransformMatrix(.., reverse=True)) of runPipeline :1155-1156,
 * which exists on the host only): uploaded once into scratch, then selected. */
int hica_order_stats_host(hica_handle h, const double *x_host, int64_t count, int32_t nk, const int64_t *k,
                          double *out, int64_t *nan_count);

#ifdef __cplusplus
}
#endif
#endif /* HICA_H */
