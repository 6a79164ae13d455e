/*
 * prb200.h — C-ABI of the B200-native PicturedRocks mutual-information hot path.
 *
 * Plain pointers and sizes only (no torch types).  Unless a function says
 * "host", every pointer is a DEVICE pointer owned by the caller; functions
 * enqueue work on `stream` (a cudaStream_t passed as void*), never allocate,
 * never synchronise, and return 0 on success (prb_last_error() explains a
 * non-zero status).  The reference has no FFI of its own: its seam for this
 * path is the flat-array numba kernels of
 *   INFOSET = picturedrocks/markers/mutualinformation/infoset.py
 * and the numpy score algebra of
 *   ITER    = picturedrocks/markers/mutualinformation/iterative.py
 *   BASE    = picturedrocks/markers/_iterative.py
 * Each entry point cites the reference lines it replaces.
 *
 * Device data layout
 *   packed CSC ("by-gene"): indptr int64[P+1]; packed uint32[nnz] with
 *     bits 0..23 = cell (row) id, bits 24..31 = bin code (1..K-1); rows
 *     ascending inside a gene (canonical CSC, INFOSET:191-196).
 *   labels: y int32[N] in 0..C-1.   classsizes int64[C].
 *   count tables: cnt int32[P][C][K-1] = #cells of class c with code v+1 in
 *     gene g (C = 1 without labels).  Zero bins are always inferred by
 *     subtraction (INFOSET:354-359, 390-391, 418-422).
 *   term table: double[N+1], T[c] = -(c/N)*log2(c/N) built on the HOST with
 *     glibc log2 (INFOSET:392-397) so entropies are bit-identical to the
 *     reference; device code only looks terms up and adds them in ascending
 *     packed-bin order.
 */
#ifndef PRB200_H
#define PRB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRB_ABI_VERSION 3
#define PRB_MAX_ROWS (1 << 24) /* cell id field of the packed CSC word */
#define PRB_MAX_CODES 256      /* K <= 256 */

typedef void *prb_stream_t;

/* ---- library ------------------------------------------------------------ */
int prb_abi_version(void);
const char *prb_last_error(void);
/* SM count and opt-in shared memory per block of the current device. */
int prb_device_info(int *sm_count, int *smem_optin_bytes);

/* ---- HOST: fp64 term table  T[c] = -(c/N)*log2(c/N), T[0] = 0 -------------
 * replaces the per-bin expression of INFOSET:137-141, 392-397, 423-427.
 * out_host: double[n_rows+1] in host memory. */
int prb_term_table(int64_t n_rows, double *out_host);

/* ---- packed CSC construction (SparseInformationSet.__init__, INFOSET:188-200;
 *      canonicalisation stays with the caller) ------------------------------ */
int prb_pack_csc(const int32_t *rows, const uint8_t *codes, int64_t nnz, uint32_t *packed,
                 prb_stream_t stream);
int prb_unpack_csc(const uint32_t *packed, int64_t nnz, int32_t *rows, uint8_t *codes,
                   prb_stream_t stream);
/* Ingestion of a caller-held canonical CSC (csrc/ingest.cu): a range of n stored entries,
 * row ids int32 | int64 (row_bytes 4 | 8), codes of any integer width (code_bytes 1|2|4|8,
 * code_signed) -> packed words.  *flags (device int32, OR-ed): 1 row outside [0, N),
 * 2 code outside [0, 255], 4 a stored zero (the reference's eliminate_zeros would drop it),
 * 8 (prb_check_columns) rows not strictly ascending inside a gene (unsorted or duplicate:
 * sum_duplicates would merge them).  Any flag: canonicalise on the host as the reference does
 * (INFOSET:191-196) and pack again. */
int prb_pack_chunk(const void *rows, int row_bytes, const void *codes, int code_bytes,
                   int code_signed, int64_t n, int64_t N, uint32_t *packed, int32_t *flags,
                   prb_stream_t stream);
int prb_check_columns(const uint32_t *packed, const int64_t *indptr, int64_t P, int32_t *flags,
                      prb_stream_t stream);
/* CSR (cells-major) -> by-gene CSC on the device, a counting transpose (csrc/ingest.cu): the
 * rows are cut into B blocks; prb_csr_count fills T int32[B][P] (entries of column c in row
 * block b; column ids outside [0, P) set flag 1); the caller scans the column totals into
 * indptr int64[P + 1]; prb_csr_fill turns T into first slots (T64 int64[B][P] scratch) and
 * scatters the entries, rows ascending inside every gene: rows_out int32[nnz], vals_out (4- or
 * 8-byte values) [nnz].  crow int64[N + 1], col int32 | int64 [nnz]; canonical CSR (the columns
 * of a row are distinct). */
int prb_csr_count(const int64_t *crow, const void *col, int col_bytes, int64_t N, int64_t P, int B,
                  int32_t *T, int32_t *flags, prb_stream_t stream);
int prb_csr_fill(const int64_t *crow, const void *col, int col_bytes, const void *vals,
                 int val_bytes, int64_t N, int64_t P, int B, int32_t *T, const int64_t *indptr,
                 int64_t *T64, int32_t *rows_out, void *vals_out, prb_stream_t stream);
/* column-major uint8 codes[P][ld] -> per-gene nnz (pass 1) and packed words (pass 2). */
int prb_dense_count_nnz(const uint8_t *codes, int64_t N, int64_t P, int64_t ld,
                        int64_t *nnz_per_gene, prb_stream_t stream);
int prb_dense_fill_csc(const uint8_t *codes, int64_t N, int64_t P, int64_t ld,
                       const int64_t *indptr, uint32_t *packed, prb_stream_t stream);

/* ---- the hot loop: per-gene joint histogram of (cell state, code) ----------
 * replaces the two-pointer merge of _sparse_entropy_wrt (INFOSET:365-391).
 * Per-cell state: uint8 (when n_states + 1 <= 256) or uint16, 0 = "not a master row"
 * (skipped; inferred by subtraction later), else 1 + joint-state id in [0, n_states).
 * The state of one TILE of cells is staged in shared memory; prb_stream_plan() gives the
 * tile size / count for (N, n_states, K-1) and prb_tile_ptrs() the per-(tile, gene)
 * sub-ranges of the packed stream: tp int64[n_tiles+1][P].
 *   hits     : int32[P][n_states*(K-1)], zero on entry; bin of a counted entry is
 *              h*(K-1) + code-1 with h = state id, or, for the direct single-gene layout
 *              (direct_C = C > 0, state id = (x_j-1)*C + y), h = y*(K-1) + x_j-1.
 *   counters : int32[n_tiles] scratch (reset by the call).
 *   state    : buffer padded to a multiple of 16 bytes. */
int prb_stream_plan(int64_t N, int64_t n_states, int K1, int64_t *tile_cells, int *n_tiles,
                    int *teams_per_cta, int *state_bytes, int64_t *smem_bytes);
int prb_tile_ptrs(const uint32_t *packed, const int64_t *indptr, int64_t P, int64_t tile_cells,
                  int n_tiles, int64_t *tp, prb_stream_t stream);
int prb_stream_hits(const uint32_t *packed, const int64_t *tp, int64_t P, int64_t nnz,
                    const void *state, int64_t N, int64_t n_states, int K1, int direct_C,
                    int32_t *hits, int32_t *counters, prb_stream_t stream);

/* ---- the hot loop for K <= 5 on the V layout (vstream.cu): 4 instructions per entry ----
 * Same job as prb_stream_hits for the direct single-gene layout — the merge of
 * INFOSET:365-391 for entropy_wrt([j]) / entropy_wrt([-1, j]) (ITER:53, 55) — and, through
 * its run lengths, the class-conditional count tables behind entropy_wrt([]) / ([-1])
 * (ITER:30, 32), on a private re-ordering of the matrix (a histogram does not depend on the
 * order in which cells are visited):
 *  1. newpos int32[N] relabels the cells so that classes are contiguous (a stable sort of
 *     the labels, by the caller; NULL = identity).  The generic packed CSC keeps the
 *     original cell ids; only the V layout and the per-cell vectors of this path (x_j, sx,
 *     the sorted labels) live in the new order.
 *  2. A host-built segment plan cuts the cells into SEGMENTS (a class, or a piece of one)
 *     and packs consecutive segments into SUPER-TILES of at most max_cells cells and
 *     max_segs segments (prb_v_tile_capacity): seg_cell0 int32[n_seg+1] ascending cell
 *     boundaries (last = N); seg_class int32[n_seg] or NULL (= all class 0);
 *     st_seg0 int32[n_st+1]: super-tile t = segments [st_seg0[t], st_seg0[t+1]).
 *  3. prb_v_count -> caller's scan -> prb_v_scatter build the V layout: every gene g
 *     becomes K-1 virtual genes (g, v), one per non-zero code v; a RUN (vg, s) holds the
 *     entries of virtual gene vg whose relabelled cell newpos[cell] lies in segment s
 *     (both calls take cellmap uint32[N], indexed by the ORIGINAL cell id: bits 0..23 =
 *     newpos[cell], bits 24..31 = segment of the cell when n_seg <= 256 — otherwise the
 *     segments come in seg16 uint16[N]; cellmap NULL = identity positions, one segment),
 *     relabelled (word = code<<24 | newpos[cell]), in arbitrary order inside a run, padded to a multiple of 32 entries (one 128-byte row); runs are
 *     stored TILE-MAJOR: by (super-tile of s, vg, s).
 *       seglen    int32[P*(K-1)][n_seg]  true run lengths
 *       run_end   int32[P*(K-1)][n_seg]  row just past each run (inclusive scan of
 *                                        ceil(seglen/32) in storage order, by the caller)
 *       tile_row0 int32[n_st+1]          first row of every super-tile's block
 *       vpacked   uint32[32 * rows]      words code<<24 | cell; the padding words of segment
 *                 s are seg_pad[s] = 0xFF000000 | ((first cell after the segment's
 *                 super-tile) mod tile span): a slot the kernel keeps at "x_j = 0"
 *  4. prb_pick_scatter (below) writes the per-cell shift code of the selected gene,
 *     sx uint8[N] (+16 bytes of padding) = 8*(x_j-1), 255 where x_j = 0 — sx is all 255
 *     between steps —, and the hsize / hs_ysum / ha tables (taken from the count table of
 *     the selected gene).
 *  5. prb_stream_v adds the joint counts to hits int32[P][C][K-1][K-1]
 *     (bin = (x_j-1)*(K-1) + x_g-1).  Work items are consecutive row ranges of a
 *     super-tile's block (large first, small last): item_off int32[n_st+1] = first item id
 *     of every super-tile; item_row0 int32[n_items + n_st] = first row of every item, the
 *     items of tile t followed by one sentinel (end of its block), i.e. item k of tile t is
 *     entry item_off[t] + t + k; item_vg0 int32[n_items] = first virtual gene with a run
 *     ending past the item's first row (a searchsorted of run_end, by the caller;
 *     prb_v_warps() = warps of one launch, to size the items).
 *     counters: int32[n_st] scratch (reset by the call unless counters_clean says the caller
 *     already did — prb_pick_scatter does); queue: device int32 scratch. */
int prb_v_tile_capacity(int64_t *max_cells, int *max_segs);
int prb_v_count(const uint32_t *packed, const int64_t *indptr, int64_t P,
                const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                int32_t *seglen, int32_t *queue, prb_stream_t stream);
int prb_v_scatter(const uint32_t *packed, const int64_t *indptr, int64_t P,
                  const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                  const int32_t *seglen,
                  const int32_t *run_end, const uint32_t *seg_pad, uint32_t *vpacked,
                  int32_t *queue, prb_stream_t stream);
/* EXPERIMENTAL "V16" layout (PRB_V16=1; staged, not yet run on hardware): rows of 64 entries
 * of 16 bits (the cell id inside its super-tile), runs padded to 64 entries, run_end / tile_row0 /
 * item_row0 in rows of 64; same arguments as prb_v_scatter / prb_stream_v. */
int prb_v_scatter16(const uint32_t *packed, const int64_t *indptr, int64_t P,
                    const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                    const int32_t *seglen, const int32_t *run_end, const uint32_t *seg_pad,
                    uint32_t *vpacked, int32_t *queue, prb_stream_t stream);
int prb_stream_v16(const uint32_t *vpacked, const int32_t *run_end, const int32_t *tile_row0,
                   int n_seg, int64_t P, int K1, int64_t v_rows, const int32_t *st_seg0,
                   const int32_t *seg_cell0, const int32_t *seg_class, int n_st,
                   int64_t max_tile_cells, const uint8_t *sx, int C, int32_t *hits,
                   int32_t *counters, const int32_t *item_off, const int32_t *item_row0,
                   const int32_t *item_vg0, int64_t n_items, int counters_clean,
                   prb_stream_t stream);
int prb_stream_v(const uint32_t *vpacked, const int32_t *run_end, const int32_t *tile_row0,
                 int n_seg, int64_t P, int K1, int64_t v_rows, const int32_t *st_seg0,
                 const int32_t *seg_cell0, const int32_t *seg_class, int n_st,
                 int64_t max_tile_cells, const uint8_t *sx, int C, int32_t *hits,
                 int32_t *counters, const int32_t *item_off, const int32_t *item_row0,
                 const int32_t *item_vg0, int64_t n_items, int counters_clean,
                 prb_stream_t stream);
int prb_v_warps(void);
/* cells per super-tile modulus (power of two) of the active kernel configuration */
int prb_v_tile_span(void);

/* ---- the hot loop on the LANE-RUN layout (vell.cu): one lane = one run, K <= 21 ----------
 * Same job and same segment plan as prb_stream_v; the default of the Python engine.  A
 * lane-run is a piece of at most CAP entries of one run (virtual gene vg, segment s); 32
 * lane-runs of one super-tile, sorted by length, form a BUNDLE stored transposed:
 * L = ceil(longest / 2) data rows whose word l holds entries 2r, 2r+1 of lane-run l (16-bit
 * cell ids inside the super-tile; missing entries = the tile's "x_j = 0" slot), and a header
 * row in a separate array bhdr uint32[n_bundles][32] (lane l: (segment index inside the
 * tile) << 24 | vg, 0xFFFFFFFF = empty lane).
 *   prb_v_count          run lengths seglen[P*(K-1)][n_seg] (as for the V layout)
 *   prb_v_scatter_flat   relabelled 16-bit entries, run-major, flat: run (vg, s) at
 *                        tmp16[run_start[vg*n_seg + s] ..) (run_start: caller's exclusive scan)
 *   (caller: split runs into lane-runs, sort by (tile, -length), cut bundles:
 *    lr_src int64 / lr_len int32 / lr_hdr uint32 [32 * n_bundles], brow0 int32[n_bundles+1],
 *    bundle_pad uint16[n_bundles])
 *   prb_ell_build        tmp16 -> ell uint32[32 * brow0[n_bundles]], bhdr
 *   prb_stream_ell       adds the joint counts to hits int32[P][C][K-1][K-1].  Work items are
 *                        ranges of whole bundles of one super-tile: item_off int32[n_st+1];
 *                        item_b0 int32[n_items + n_st] = first bundle of every item, the items
 *                        of tile t followed by one sentinel (first bundle of tile t+1).
 *                        PV = P*(K-1); a header word that names no run of its tile is never
 *                        written through: it is counted in dbg[0] (device int32[16], may be
 *                        NULL; dbg[1..15] describe the first one).
 * prb_v_check_smem_base(): 0 when dynamic shared memory starts at the window offset the
 * V-layout kernels assume (they address the cell tile by the bare cell index). */
int prb_v_scatter_flat(const uint32_t *packed, const int64_t *indptr, int64_t P,
                       const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                       const int32_t *seglen, const int64_t *run_start, uint16_t *tmp16,
                       int32_t *queue, prb_stream_t stream);
int prb_ell_build(const uint16_t *tmp16, const int64_t *lr_src, const int32_t *lr_len,
                  const uint32_t *lr_hdr, const int32_t *brow0, const uint16_t *bundle_pad,
                  int64_t n_bundles, uint32_t *ell, uint32_t *bhdr, int hdr_in_stream,
                  prb_stream_t stream);
int prb_stream_ell(const uint32_t *ell, const uint32_t *bhdr, const int32_t *brow0,
                   int64_t n_bundles, int K1,
                   const int32_t *st_seg0, const int32_t *seg_cell0, const int32_t *seg_class,
                   int n_st, int64_t max_tile_cells, const uint8_t *sx, int C, int32_t *hits,
                   int32_t *counters, const int32_t *item_off, const int32_t *item_b0,
                   int64_t n_items, int counters_clean, int64_t PV, int32_t *dbg,
                   prb_stream_t stream);
int prb_ell_warps(void);
int prb_v_check_smem_base(void);

/* ---- the hot loop on the STREAMED lane-run layout (vell2.cu) — the default ------------------
 * Same lane-runs and bundles as above, laid out so that a warp needs nothing but its ring of
 * bulk copies:
 *  - super-tiles hold at most `tile_cells` cells (prb_ell2_geometry) and tile t owns the
 *    positions [t * 32768, t * 32768 + cells) of a PADDED cell space: pos int32[N] replaces
 *    newpos in prb_pick_scatter / prb_finalize_update / prb_v_unscatter, sx has n_st * 32768
 *    slots, slot 32767 of every tile stays "x_j = 0" (what padding entries point at), and a
 *    16-bit entry = pos & 0xFFFF addresses a 64 KB window of two consecutive tiles;
 *  - prb_ell_build(.., hdr_in_stream = 1): brow0 counts STREAM rows — per bundle one header row
 *    followed by its L data rows; header word of lane l = bit 31: bit l of L, bit 30: the
 *    lane-run is a piece of a longer run, bits id_shift..29: the run's slot in a gene's block
 *    of hits (its segment or its class), bits 0..id_shift-1: virtual gene (all ones = empty
 *    lane); bhdr unused;
 *  - the stream is cut on the host into one chunk of bundles per CTA, a chunk into phases
 *    (the bundles of at most two consecutive tiles) and a phase into at most `max_items` work
 *    items of decreasing size: item_row0 int32[n_items + 1] (first stream row of every item, the
 *    items of all phases back to back), phase int32[n_phases][4] = {first tile, tiles (1 or 2;
 *    | 256: the items are claimed last to first), first item, items}, cta_ph0 int32[n_ctas + 1] = the phases of every CTA's own chunk; the
 *    phases [pool0, pool1) form a pool that the CTAs draw from once their chunk is done,
 *    through the device counter *pool_ctr (zero on entry: prb_pick_scatter resets it).
 * prb_stream_ell2, vmajor = 0: adds the joint counts to hits int32[P][n_slot][K-1][K-1];
 * vmajor = 1: hits uint16[P][n_slot][K-1][4 * ceil((K-1)/4)] with one slot per SEGMENT (a count
 * is at most 32 752), the counts of x_j - 1 = 4q, 4q+2, 4q+1, 4q+3 in the last dimension, all
 * zero on entry — a lane-run that holds a whole run STORES its counts (one 8-byte write per
 * four values of x_j), only the pieces of longer runs are added (32-bit REDs on pairs of
 * counts: they cannot carry).  The slot of a run is the id field of its header when
 * n_slot > 1, else 0.  PV = P*(K-1) < 2^id_shift - 1.
 * timing: NULL, or device int64[n_items][2] that receives the start / end time (ns) of every
 * work item (tuning runs, K <= 5). */
int prb_ell2_geometry(int *tile_cells, int *ctas, int *warps, int *max_items);
int prb_stream_ell2(const uint32_t *ell, const int32_t *item_row0, const int32_t *phase,
                    const int32_t *cta_ph0, int n_ctas, int pool0, int pool1, int32_t *pool_ctr,
                    const uint8_t *sx, int K1, int n_slot,
                    int id_shift, int32_t *hits, int vmajor, int64_t PV, int64_t *timing,
                    prb_stream_t stream);

/* ---- joint-state build ("master column", _sparse_make_master_col INFOSET:431-442) */
/* Columns are addressed through col_begin / col_end int64[P]: first / past-the-last word of
 * every gene in `packed` (indptr[:-1] / indptr[1:] of a contiguous CSC; multi-GPU: the slots
 * of the replicated copy of all ranks' shards, gene ids global).
 * dense uint8 x_j[N] (zeroed by caller) <- codes of gene *gene_ptr (device int32), written
 * at newpos[cell] (NULL = at cell). */
int prb_scatter_gene_u8(const uint32_t *packed, const int64_t *col_begin, const int64_t *col_end,
                        const int32_t *gene_ptr, int64_t P, const int32_t *newpos, uint8_t *xj,
                        prb_stream_t stream);
/* mpacked[row] |= code << shift_bits for every stored entry of gene `gene`. */
int prb_scatter_gene_or32(const uint32_t *packed, const int64_t *col_begin,
                          const int64_t *col_end, int64_t gene, int shift_bits, uint32_t *mpacked,
                          prb_stream_t stream);
/* direct single-column state: state = 1 + (x_j-1)*C + y for x_j > 0 (y = 0 without
 * labels).  Writes state[N] (state_bytes 1|2), hsize int32[C*(K-1)] indexed
 * y*(K-1)+x_j-1 (zeroed here), hs_begin int32[C+1], hs_y int32[C*(K-1)],
 * hs_ysum int32[C], ha int32[K-1]. */
int prb_state_direct(const uint8_t *xj, const int32_t *y, int64_t N, int C, int K, void *state,
                     int state_bytes, int32_t *hsize, int32_t *hs_begin, int32_t *hs_y,
                     int32_t *hs_ysum, int32_t *ha, prb_stream_t stream);
/* generic multi-column state through an order-preserving presence-table compaction
 * (the "sort/unique relabel"): key = y << mbits | mpacked.  table int32[C << mbits]
 * scratch; after the call it maps key -> rank.  n_hs_out: device int32.
 * prb_state_assign then writes state = 1 + rank (0 where mpacked == 0). */
int prb_state_table(const uint32_t *mpacked, const int32_t *y, int64_t N, int C, int mbits,
                    int32_t *table, int32_t *hsize, int32_t *hs_begin, int32_t *hs_y,
                    int32_t *hs_ysum, int32_t *n_hs_out, int64_t hsize_cap, prb_stream_t stream);
int prb_state_assign(const uint32_t *mpacked, const int32_t *y, int64_t N, int mbits,
                     const int32_t *table, void *state, int state_bytes, prb_stream_t stream);

/* ---- entropies from histograms (INFOSET:392-397), ordered fp64 sum ----------
 * out_full[g]  = H(y, master, x_g)   (bins ascending (y, master, code))
 * out_marg[g]  = H(master, x_g)      (only for the direct layout; may be NULL)
 * hits may be NULL (no master column: gives H(y,x_g) and H(x_g)).
 * hits entries are reset to zero as they are consumed. */
int prb_finalize(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                 const int32_t *hs_begin, const int32_t *hs_y, const int32_t *hsize,
                 const int32_t *hs_ysum, const int32_t *ha, const int32_t *n_hs_ptr,
                 int64_t n_hs_cap, int C, int K, const double *table, int64_t N, int64_t P,
                 double *out_full, double *out_marg, prb_stream_t stream);
/* Same result for the DIRECT single-gene layout (state id = y*(K-1) + x_j - 1, as
 * written by prb_state_direct), one thread per gene; hits int32[P][C][K-1][K-1] or
 * NULL (then out_full = H(y, x_g), out_marg = H(x_g)).  vmajor != 0 (K <= 5): hits is the
 * table prb_stream_ell2 stores, uint16[P][n_slot][K-1][4] with the counts of x_j - 1 = 0, 2,
 * 1, 3 in the last dimension; the slots of class y are [cls_slot0[y], cls_slot0[y+1]) — one
 * per segment of the class, added up here — or the single slot y when cls_slot0 is NULL. */
int prb_finalize_direct(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                        const int32_t *hsize, const int32_t *hs_ysum, const int32_t *ha, int C,
                        int K, const double *table, int64_t N, int64_t P, double *out_full,
                        double *out_marg, int vmajor, const int32_t *cls_slot0, int n_slot,
                        prb_stream_t stream);
/* scalar entropy of a per-cell key vector through a count table (INFOSET:218-249,
 * 413-428): table int32[n_keys] scratch; out: device double. */
int prb_entropy_keys(const uint32_t *mpacked, const int32_t *y, int64_t N, int mbits,
                     int64_t n_keys, int32_t *table, const double *term_table, double *out,
                     prb_stream_t stream);

/* ---- selector algebra + argmax on device (ITER:24-134, BASE:53-66) --------- */
enum {
    PRB_SEL_CIFE = 0,
    PRB_SEL_JMI = 1,
    PRB_SEL_CIFE_UNSUP = 2
};
/* base = (Hx + Hy) - Hyx  (ITER:29-33) */
int prb_mi_base(const double *Hx, double Hy, const double *Hyx, int64_t P, double *base,
                prb_stream_t stream);
/* one add()/remove() update (ITER:48-72, 76-100, 112-134).
 *   delta = (((base + Hj) - Hij) - Hyj) + Hyij   (supervised)
 *   delta = (base + Hj) - Hij                     (CIFEUnsup)
 * Hj/Hyj are looked up at *gene_ptr in the replicated vectors Hx_all/Hyx_all.
 * n_sel_ptr: device int32 = len(S) AFTER the add/remove.
 * selected: uint8[P] mask of S (local genes).  Also produces per-block argmax
 * candidates (cand_score/cand_idx, one per block of 256 genes; idx is global =
 * gene_lo + local). */
int prb_selector_update(int kind, int is_remove, const double *base, double *penalty,
                        double *score, const uint8_t *selected, const double *Hij,
                        const double *Hyij, const double *Hx_all, const double *Hyx_all,
                        const int32_t *gene_ptr, const int32_t *n_sel_ptr, int64_t P,
                        int32_t gene_lo, double *cand_score, int64_t *cand_idx,
                        prb_stream_t stream);
/* per-block argmax candidates of a score vector (first maximum wins; NaN counts
 * as maximal, like np.argmax). */
int prb_argmax_blocks(const double *score, int64_t P, int32_t gene_lo, double *cand_score,
                      int64_t *cand_idx, prb_stream_t stream);
/* reduce n_cand candidates to one (score,idx); ties -> lowest idx.  Candidate c is
 * read at cand_score[c*stride], cand_idx[c*stride] (stride 2 = interleaved 16-byte
 * (score, idx) records, the layout of the multi-GPU all-gather). */
int prb_argmax_final(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                     int64_t stride, double *best_score, int64_t *best_idx,
                     prb_stream_t stream);
/* prb_argmax_final followed by prb_commit_add, one launch. */
int prb_argmax_commit(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                      int64_t stride, double *best_score, int64_t *best_idx, int32_t *gene_ptr,
                      int32_t *S, int32_t *n_sel_ptr, uint8_t *selected, int32_t gene_lo, int64_t P,
                      prb_stream_t stream);
/* commit a winner: gene_ptr <- idx, S[n_sel++] <- idx, selected[idx-gene_lo] <- 1,
 * score[idx-gene_lo] untouched (the update masks it).  idx is read from best_idx. */
int prb_commit_add(const int64_t *best_idx, int32_t *gene_ptr, int32_t *S, int32_t *n_sel_ptr,
                   uint8_t *selected, int32_t gene_lo, int64_t P, prb_stream_t stream);

/* ---- the greedy step on the V-layout path in three launches (csrc/step.cu, finalize.cu) ----
 * prb_pick_scatter -> prb_stream_v[16] (counters_clean = 1) -> prb_finalize_update.
 *
 * prb_pick_scatter: np.argmax + add()'s bookkeeping (BASE:64-66, ITER:48-49) and the master
 * column of the winner (INFOSET:289, 431-442).  records double[n_rec][2] = (score, index as
 * int64 bits) candidates, one per rank (all-gathered) or this rank's own; reduced by every
 * block with np.argmax's rule.  n_rec = 0: the gene already in *gene_ptr (add(ind) /
 * remove(ind) / entropy_wrt([j]) — nothing is committed).  Commit: *gene_ptr <- j,
 * S[(*n_sel_ptr)++] <- j, selected[j - gene_lo] <- 1 when local.  cnt_all int32[P_total][C][K-1]
 * (count table of ALL genes) gives hsize[y][a] = cnt_all[j][y][a], hs_ysum[y], ha[a] (and the
 * hs_begin / hs_y view of the same layout that prb_finalize reads).
 * counters[n_counters] <- 0.  Column j (col_packed / col_begin / col_end, global gene ids)
 * is scattered into sx (all 255 on entry) at newpos[cell]. */
int prb_pick_scatter(const double *records, int n_rec, int32_t *gene_ptr, int32_t *S,
                     int32_t *n_sel_ptr, uint8_t *selected, int32_t gene_lo, int64_t P_local,
                     const uint32_t *col_packed, const int64_t *col_begin, const int64_t *col_end,
                     const int32_t *newpos, uint8_t *sx, const int32_t *cnt_all, int C, int K,
                     int32_t *hsize, int32_t *hs_begin, int32_t *hs_y, int32_t *hs_ysum,
                     int32_t *ha, int32_t *counters, int n_counters, const uint64_t *peer_local,
                     int n_peers, const int32_t *xtag, int32_t *peer_err, prb_stream_t stream);
/* The per-step exchange of the gene-sharded loop over NVLink peer memory (peer.cu) instead of an
 * NCCL all-gather: peer_bufs = device array of W pointers, entry r = rank r's record buffer
 * uint64[2][W][4] as mapped into THIS process (CUDA IPC); the call stores this rank's record
 * best double[2] = (score, index bits) into slot [tag & 1][rank] of every buffer, then the tag
 * (release, system scope), tag = ++*xtag.  prb_pick_scatter with peer_local = this rank's own
 * buffer (records may be NULL) and n_peers = W waits until all W slots of the half *xtag & 1
 * carry the tag *xtag — also when it does not pick (n_rec = 0), so that no rank runs more than
 * one exchange ahead — and reduces them (n_rec = W); *peer_err is set when a peer has not
 * answered within 4 s. */
int prb_peer_share(const double *best, const uint64_t *const *peer_bufs, int W, int rank,
                   int32_t *xtag, prb_stream_t stream);
/* prb_finalize_direct + prb_selector_update + the reduction of the block candidates to ONE
 * record best double[2] = (score, index bits) of this rank (done by the last block; ticket:
 * device int32, zero on entry and exit; cand_score / cand_idx: scratch of
 * prb_finalize_update_blocks(P, C) entries) + the reset of column *gene_ptr of sx to 255.
 * Supervised kinds use H(y,x_j,x_g) and H(x_j,x_g); PRB_SEL_CIFE_UNSUP (C = 1) H(x_j,x_g).
 * peer_bufs != NULL: the last block also does prb_peer_share(best, peer_bufs, W, rank, xtag). */
int64_t prb_finalize_update_blocks(int64_t P, int C);
int prb_finalize_update(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                        const int32_t *hsize, const int32_t *hs_ysum, const int32_t *ha, int C,
                        int K, const double *table, int64_t N, int64_t P, int kind, int is_remove,
                        const double *base, double *penalty, double *score,
                        const uint8_t *selected, const double *Hx_all, const double *Hyx_all,
                        const int32_t *gene_ptr, const int32_t *n_sel_ptr, int32_t gene_lo,
                        double *cand_score, int64_t *cand_idx, double *best, int32_t *ticket,
                        const uint32_t *col_packed, const int64_t *col_begin,
                        const int64_t *col_end, const int32_t *newpos, uint8_t *sx, int vmajor,
                        const int32_t *cls_slot0, int n_slot, const uint64_t *const *peer_bufs,
                        int W, int rank, int32_t *xtag, prb_stream_t stream);
/* The same tail for 5 < K <= 21 (finalize_wide.cu) on the v-major table uint16
 * hits[P][C][K-1][4*ceil((K-1)/4)] written by prb_stream_ell2(vmajor = 1, n_slot = C) — every
 * class must be ONE segment of the layout —, in two kernels: k_wide_terms turns every gene's
 * block into its ordered list of fp64 terms (terms double[P][prb_wide_terms_stride(C, K)],
 * n_terms int32[P]; the K*K terms of H(x_j, x_g) in terms2 double[P][K*K] for the supervised
 * kinds) and leaves the table all zero; k_wide_chain adds them one thread per gene and runs
 * the fused tail.  cand_score / cand_idx need prb_wide_chain_blocks(P) entries.
 * prb_wide_supported(C, K): the per-gene scratch fits in shared memory. */
int64_t prb_wide_terms_stride(int C, int K);
int prb_wide_supported(int C, int K);
int64_t prb_wide_chain_blocks(int64_t P);
int prb_finalize_update_wide(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                             const int32_t *hsize, const int32_t *hs_ysum, const int32_t *ha,
                             int C, int K, const double *table, int64_t N, int64_t P,
                             double *terms, double *terms2, int32_t *n_terms, int kind,
                             int is_remove, const double *base, double *penalty, double *score,
                             const uint8_t *selected, const double *Hx_all, const double *Hyx_all,
                             const int32_t *gene_ptr, const int32_t *n_sel_ptr, int32_t gene_lo,
                             double *cand_score, int64_t *cand_idx, double *best, int32_t *ticket,
                             const uint32_t *col_packed, const int64_t *col_begin,
                             const int64_t *col_end, const int32_t *newpos, uint8_t *sx,
                             const uint64_t *const *peer_bufs, int W, int rank, int32_t *xtag,
                             prb_stream_t stream);
/* sx <- 255 at the stored entries of gene *gene_ptr (after a stand-alone streaming pass). */
int prb_v_unscatter(const int32_t *gene_ptr, const uint32_t *col_packed, const int64_t *col_begin,
                    const int64_t *col_end, const int32_t *newpos, uint8_t *sx,
                    prb_stream_t stream);
/* block candidates of prb_argmax_blocks -> best double[2] (score, index bits). */
int prb_best_record(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                    double *best, prb_stream_t stream);

/* ---- discretiser (quantile_discretize, INFOSET:50-83) ---------------------
 * dtype: 0 = float32, 1 = float64, 2 = int64 (edges float64 for int64 input).
 * Pipeline: row-major X[N][P] --transpose--> Xt[P][N] --sort_columns--> Xs[P][N]
 *           --qd_edges--> edges[P][k-1], n_edges[P] --qd_digitize(Xt)--> codes.
 * Edges follow numpy-2.x percentile(method="linear") arithmetic in the input
 * dtype (float32 stays float32); codes = #edges strictly below x
 * (np.digitize(right=True)). */
/* row-major [N][P] -> column-major [P][N] transpose of 4- or 8-byte elements. */
int prb_transpose(const void *X, int elem_bytes, int64_t N, int64_t P, void *out,
                  prb_stream_t stream);
/* ascending sort of every gene segment of a column-major [P][N] array (N*P < 2^31);
 * workspace is caller-provided device scratch of prb_sort_columns_workspace() bytes. */
int prb_sort_columns_workspace(int dtype, int64_t N, int64_t P, int64_t *bytes);
int prb_sort_columns(const void *cols_in, void *cols_out, int dtype, int64_t N, int64_t P,
                     void *workspace, int64_t workspace_bytes, prb_stream_t stream);
int prb_qd_edges(const void *sorted, int dtype, int64_t N, int64_t P, int k, void *edges,
                 int32_t *n_edges, prb_stream_t stream);
/* Xt column-major [P][N] (unsorted); codes column-major uint8 [P][ld]. */
int prb_qd_digitize(const void *Xt, int dtype, int64_t N, int64_t P, int k, const void *edges,
                    const int32_t *n_edges, uint8_t *codes, int64_t ld, prb_stream_t stream);

/* Sparse input (scipy CSR/CSC in the reference; INFOSET:66-67 densifies it, so the implicit
 * zeros take part in the quantiles): by-gene CSC of RAW values on the device, rows ascending
 * inside a gene.  Per block of genes: prb_sort_segments (ascending sort of every gene's stored
 * values; offsets int64[n_segs+1] relative to vals_in; < 2^31 values per call) ->
 * prb_qd_edges_sparse (edges as prb_qd_edges, the column being the sorted stored values plus
 * N - nnz zeros; indptr points at the block's first gene, sorted_vals at its first value;
 * zero_code[g] = code of the implicit zeros, 0 for non-negative data — the coded matrix is
 * sparse only then).  Over all genes: prb_qd_count_sparse -> caller's scan ->
 * prb_qd_fill_sparse digitise the stored values and emit the packed CSC, dropping the
 * entries whose code is 0 (the reference's eliminate_zeros, INFOSET:195). */
int prb_sort_segments_workspace(int dtype, int64_t n_items, int64_t n_segs, int64_t *bytes);
int prb_sort_segments(const void *vals_in, void *vals_out, int dtype, int64_t n_items,
                      int64_t n_segs, const int64_t *offsets, void *workspace,
                      int64_t workspace_bytes, prb_stream_t stream);
int prb_qd_edges_sparse(const void *sorted_vals, const int64_t *indptr, int dtype, int64_t N,
                        int64_t P, int k, void *edges, int32_t *n_edges, int32_t *zero_code,
                        prb_stream_t stream);
int prb_qd_count_sparse(const int32_t *rows, const void *vals, const int64_t *indptr, int dtype,
                        int64_t P, int k, const void *edges, const int32_t *n_edges,
                        int64_t *nnz_per_gene, prb_stream_t stream);
int prb_qd_fill_sparse(const int32_t *rows, const void *vals, const int64_t *indptr, int dtype,
                       int64_t P, int k, const void *edges, const int32_t *n_edges,
                       const int64_t *out_indptr, uint32_t *packed, prb_stream_t stream);

/* ---- synthetic discretised matrices (bench/test utility; counter-based RNG,
 *      reproduced on the CPU by picturedrocks_b200/synth.py) -----------------
 * cum: uint32[4][K-2] cumulative 16-bit thresholds of the code distribution. */
int prb_synth_labels(int32_t *y, int64_t N, int C, uint64_t seed, prb_stream_t stream);
int prb_synth_count(const int32_t *y, int64_t N, int64_t gene_lo, int64_t P, int C, int K,
                    uint32_t mean_density_q24, uint64_t seed, const uint32_t *cum,
                    int64_t *nnz_per_gene, prb_stream_t stream);
int prb_synth_fill(const int32_t *y, int64_t N, int64_t gene_lo, int64_t P, int C, int K,
                   uint32_t mean_density_q24, uint64_t seed, const uint32_t *cum,
                   const int64_t *indptr, uint32_t *packed, prb_stream_t stream);

/* ---- HOST-buffer drop-in for the reference's njit seam ---------------------
 * Same arguments and meaning as
 *   _sparse_entropy_wrt(dindices, dindptr, ddata, mindices, mdata, n_rows,
 *                       n_feats, n_mcols, shift, ybits, y, classsizes)
 * (INFOSET:306-349).  All pointers are HOST pointers; the call copies inputs to
 * the current device, runs the kernels above and copies float64[n_feats] back.
 * Allocates and synchronises internally (it is the synchronous reference
 * signature).  m_nnz = len(mindices); n_classes = len(classsizes) (0 if no y). */
int prb_sparse_entropy_wrt_host(const int32_t *dindices, const int64_t *dindptr,
                                const int64_t *ddata, const int64_t *mindices,
                                const int64_t *mdata, int64_t m_nnz, int64_t n_rows,
                                int64_t n_feats, int64_t n_mcols, int64_t shift, int64_t ybits,
                                const int64_t *y, const double *classsizes, int64_t n_classes,
                                double *out);

/* Same for the scalar kernel
 *   _sparse_entropy(indices, data, n_rows, max_val)          (INFOSET:413-428)
 * The reference uses `indices` only for its length: pass nnz = len(indices) and the packed
 * values `data` (HOST pointer, int64, each in [0, max_val]); *out (host) receives the
 * entropy.  Bit-identical to the reference (same bins, same ascending order). */
int prb_sparse_entropy_host(const int64_t *data, int64_t nnz, int64_t n_rows, int64_t max_val,
                            double *out);

#ifdef __cplusplus
}
#endif
#endif /* PRB200_H */
