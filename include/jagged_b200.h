/*
 * jagged_b200.h -- C ABI of libjagged_sm100a.so: the awkward0 JaggedArray hot path on B200 (sm_100a).
 *
 * The reference (scikit-hep/awkward-0.x 0.15.5) is pure Python on numpy and has NO native interface; its
 * "kernels" are the numpy primitives called from awkward0/array/jagged.py.  Each entry point below replaces
 * one of those call sites (cited as jagged.py:LINE) and is what a ctypes binding inside the reference's own
 * JaggedArray would call (INTEGRATION.md shows that binding).
 *
 * Conventions
 *   - every pointer is a RAW DEVICE pointer unless the parameter is documented as host;
 *   - sizes and indices are int64_t (the reference's INDEXTYPE, base.py:44);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, re-entrant and
 *     keeps no state between calls (the only global is the thread-local error string);
 *   - the library never allocates or frees device memory: scratch comes from the caller (`ws`, `ws_bytes`,
 *     size from jg_workspace_bytes), results go to caller-allocated buffers;
 *   - return value: 0 = launched, <0 = argument / CUDA error (text in jg_last_error());
 *   - data-dependent conditions (negative counts, non-monotone offsets, out-of-bounds jagged index, ...) are
 *     OR-ed into a caller-provided device word `status` (JG_ST_* bits); the Python host reads it together
 *     with the scan total and raises the reference's exception type and message.
 *   - `dtype` arguments use the JG_* codes below; bool content is JG_B8 (one byte, 0/1).
 */
#ifndef JAGGED_B200_H
#define JAGGED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

/* ---- dtype codes -------------------------------------------------------------------------------------- */
enum {
    JG_F32 = 0, JG_F64 = 1, JG_I32 = 2, JG_I64 = 3, JG_B8 = 4, JG_U8 = 5,
    JG_I8 = 6, JG_I16 = 7, JG_U16 = 8, JG_U32 = 9, JG_U64 = 10
};

/* ---- status bits (device word written by kernels) ------------------------------------------------------ */
enum {
    JG_ST_NEGATIVE_COUNT  = 1,   /* fromcounts: "counts must be a non-negative array"        jagged.py:160 */
    JG_ST_NEGATIVE_OFFSET = 2,   /* "offsets must be a non-negative array"                   jagged.py:126 */
    JG_ST_NONMONOTONE     = 4,   /* "offsets must be monatonically increasing"               jagged.py:474 */
    JG_ST_BEYOND_CONTENT  = 8,   /* "maximum offset {0} is beyond the length of the content" jagged.py:476 */
    JG_ST_INDEX_OOB       = 16,  /* "jagged array used as index contains out-of-bounds"      jagged.py:555 */
    JG_ST_COUNTS_MISMATCH = 32,  /* "cannot fit contents of JaggedArray into the given ..."  jagged.py:911 */
    JG_ST_NOT_OFFSETS     = 64,  /* starts[1:] != stops[:-1]                                 jagged.py:358 */
    JG_ST_STOPS_LT_STARTS = 128, /* "stops must be greater than or equal to starts"          jagged.py:501 */
    JG_ST_ARG_ALLNAN      = 256, /* a non-empty event holds only NaN: redo argmin/argmax with the exact path   */
    JG_ST_SORT_LONG       = 512  /* jg_argsort left events to jg_argsort_long (their indices are in long_list)  */
};

/* ---- reducers (base.py:193-215) -------------------------------------------------------------------------- */
enum { JG_SUM = 0, JG_PROD = 1, JG_MIN = 2, JG_MAX = 3, JG_ANY = 4, JG_ALL = 5, JG_COUNT_NONZERO = 6 };

/* ---- elementwise ops (numpy ufuncs reached through __array_ufunc__, jagged.py:944-1051) ------------------- */
enum {
    JG_OP_ADD = 0, JG_OP_SUB = 1, JG_OP_MUL = 2, JG_OP_DIV = 3, JG_OP_FLOORDIV = 4, JG_OP_MOD = 5,
    JG_OP_POW = 6, JG_OP_MAXIMUM = 7, JG_OP_MINIMUM = 8, JG_OP_ARCTAN2 = 9, JG_OP_HYPOT = 10,
    JG_OP_GT = 16, JG_OP_GE = 17, JG_OP_LT = 18, JG_OP_LE = 19, JG_OP_EQ = 20, JG_OP_NE = 21,
    JG_OP_AND = 24, JG_OP_OR = 25, JG_OP_XOR = 26,
    JG_OP_NEG = 32, JG_OP_ABS = 33, JG_OP_SQRT = 34, JG_OP_EXP = 35, JG_OP_LOG = 36, JG_OP_SIN = 37,
    JG_OP_COS = 38, JG_OP_TAN = 39, JG_OP_SINH = 40, JG_OP_COSH = 41, JG_OP_TANH = 42, JG_OP_SQUARE = 43,
    JG_OP_NOT = 44, JG_OP_ISNAN = 45, JG_OP_CAST = 46, JG_OP_ISFINITE = 47, JG_OP_ARCSINH = 48,
    JG_OP_FLOOR = 49, JG_OP_CEIL = 50, JG_OP_RINT = 51, JG_OP_SIGN = 52, JG_OP_LOG10 = 53, JG_OP_EXP2 = 54,
    JG_OP_ARCSIN = 55, JG_OP_ARCCOS = 56, JG_OP_ARCTAN = 57, JG_OP_ISINF = 58, JG_OP_LOG2 = 59,
    JG_OP_LOG1P = 60, JG_OP_EXPM1 = 61, JG_OP_RECIPROCAL = 62
};

/* how the second operand of a binary op is addressed */
enum {
    JG_RHS_ARRAY  = 0,   /* same length as lhs                                            */
    JG_RHS_SCALAR = 1,   /* one value (taken from `scalar_f64` / `scalar_i64`)             */
    JG_RHS_PARENT = 2    /* flat[N] broadcast through the events: rhs[parent(j)]  jagged.py:1003-1037 */
};

/* ---- combinatorics kinds ---------------------------------------------------------------------------------- */
enum {
    JG_COMB_PAIRS = 0,      /* i <= j, lexicographic             jagged.py:1070-1091 */
    JG_COMB_DISTINCTS = 1,  /* i <  j, lexicographic             jagged.py:1093-1126 */
    JG_COMB_CHOOSE = 2,     /* k = 2..5, colexicographic         jagged.py:1128-1221 */
    JG_COMB_CROSS = 3       /* rectangular, left index slow      jagged.py:1302-1325 */
};
/* OR-ed into `kind` of jg_comb_fill / jg_comb_gather2 / jg_comb_gather_multi / jg_comb_eval: a HINT that events hold
 * more than a handful of combinations each (the caller knows the total and the number of events).  The result is the
 * same either way; with the hint the kernels built for long events run (direct tables: owner table over output
 * chunks + per-event info + combination look-up table) instead of those tuned for dimuon-like events.              */
#define JG_COMB_LONG 0x100
/* OR-ed into `op` of jg_segreduce and into `is_min` of jg_segargminmax_dense / _dense_tiles: a HINT that the mean event
 * holds 7 - 14 items (the caller knows content length and number of events).  The result is the same either way; with
 * the hint 4-byte contents are staged 32 KB per 512-event tile instead of 16 KB, so such arrays stay on the
 * thread-per-event kernels instead of taking the long-tile path (ignored for other item sizes).                     */
#define JG_WIDE_STAGE 0x100
/* the same for a mean of 14 - 28 items: 64 KB per tile (dynamic shared memory, 3 resident CTAs per SM)               */
#define JG_WIDER_STAGE 0x200

/* ---- library ------------------------------------------------------------------------------------------------ */
int         jg_version(void);
const char *jg_last_error(void);                         /* thread-local, host string */
int         jg_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* scratch bytes sufficient for any call on `n` events (scan descriptors + ticket); CUB-style two-phase API */
size_t      jg_workspace_bytes(int64_t n);

/* ---- offsets / parents machinery --------------------------------------------------------------------------- */
/* jagged.py:42-47  counts2offsets: offsets[0]=0, offsets[1:]=cumsum(counts).  Single-pass decoupled-lookback
 * scan.  counts_dtype in {JG_I32, JG_I64, JG_U8, JG_B8}; *total (device int64) receives offsets[n];
 * JG_ST_NEGATIVE_COUNT is raised in *status for negative counts (jagged.py:160).                              */
int jg_counts2offsets(const void *counts, int counts_dtype, int64_t n, int64_t *offsets, int64_t *total,
                      int32_t *status, void *ws, size_t ws_bytes, void *stream);
/* The same scan that also reports, for every tile of 512 consecutive events, how many events are non-empty
 * (tile_nonempty: ceil(n / 512) entries; int32 / int64 counts).  jg_segargminmax_dense_tiles takes that array and
 * skips its own 8N-byte counting pass: the reference's argmin / argmax (jagged.py:1581-1601) on an array built by
 * fromcounts (jagged.py:155-166) then moves exactly its algorithmic bytes.                                          */
int jg_counts2offsets_tiles(const void *counts, int counts_dtype, int64_t n, int64_t *offsets, int64_t *tile_nonempty,
                            int64_t *total, int32_t *status, void *ws, size_t ws_bytes, void *stream);

/* jagged.py:383-388  counts = stops - starts (offsets[1:] - offsets[:-1]) */
int jg_offsets2counts(const int64_t *offsets, int64_t n, int64_t *counts, void *stream);
/* general layout: counts[i] = stops[i] - starts[i] */
int jg_startsstops2counts(const int64_t *starts, const int64_t *stops, int64_t n, int64_t *counts, void *stream);
/* jagged.py:120-127,469-477  _valid(): non-negative, monotone, max <= content_len.
 * info (device int64[4]) = {offsets[0], offsets[n], max count, number of non-empty events}.                    */
int jg_validate_offsets(const int64_t *offsets, int64_t n, int64_t content_len, int64_t *info,
                        int32_t *status, void *stream);
/* jagged.py:358,495-502  general starts/stops: stops >= starts, bounds, and whether starts[1:] == stops[:-1].
 * info (device int64[4]) = {min start of non-empty, max stop, max count, number of non-empty events}.          */
int jg_validate_startsstops(const int64_t *starts, const int64_t *stops, int64_t n, int64_t content_len,
                            int64_t *info, int32_t *status, void *stream);
/* jagged.py:49-54  offsets2parents: parents[j] = i for offsets[i] <= j < offsets[i+1]; -1 below offsets[0].
 * parents has offsets[n] entries (the caller knows offsets[n] from jg_validate_offsets / the scan total).      */
int jg_offsets2parents(const int64_t *offsets, int64_t n, int64_t *parents, void *stream);
/* jagged.py:430-437  localindex content: out[j - offsets[0]] = j - offsets[parent(j)]                          */
int jg_localindex(const int64_t *offsets, int64_t n, int64_t *out, void *stream);
/* jagged.py:56-71  startsstops2parents for a non-dense layout: out (length out_len) pre-filled with -1          */
int jg_startsstops2parents(const int64_t *starts, const int64_t *stops, int64_t n, int64_t *out,
                           int64_t out_len, void *stream);
/* jagged.py:883-942 / 1386-1398  compact(): out[dense_offsets[i] + k] = content[starts[i] + k]                  */
int jg_compact_gather(const void *content, int itemsize, const int64_t *starts, const int64_t *dense_offsets,
                      int64_t n, void *out, void *stream);

/* ---- segmented reductions ------------------------------------------------------------------------------------ */
/* jagged.py:1459-1565 via base.py:193-215.  op in JG_SUM..JG_ALL.  NaN -> identity, empty -> identity.
 * Output dtype follows numpy: sum/prod of B8/I32 -> I64 (U8 -> U64), min/max keep the dtype (B8 -> F64),
 * any/all -> B8; the caller passes the matching out buffer.                                                      */
int jg_segreduce(const void *content, int dtype, const int64_t *offsets, int64_t n, int op, void *out,
                 int64_t content_len, void *stream);
/* jagged.py:1567-1601  argmin/argmax, logical value: best[i] = local index of the first extreme non-NaN element,
 * or -1 for an event with none.  Optionally also writes the extreme value itself (same dtype as content).        */
int jg_segargminmax(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min,
                    int64_t *best, void *best_value, void *stream);
/* the dense result in one optimistic pass: out_offsets[n+1] (scanned 0/1 counts), out_index[<= n] (allocate n),
 * *total = number of events with a result.  Assumes every non-empty event has a result; when a non-empty event
 * holds only NaN, JG_ST_ARG_ALLNAN is raised in *status, the outputs are not valid and the caller must use
 * jg_segargminmax + jg_compact_nonneg instead.                                                                    */
int jg_segargminmax_dense(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min,
                          int64_t *out_offsets, int64_t *out_index, void *out_value, int64_t *total, int32_t *status,
                          void *ws, size_t ws_bytes, void *stream);
int jg_segargminmax_dense_tiles(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min,
                                const int64_t *tile_nonempty /* from jg_counts2offsets_tiles, or NULL */,
                                int64_t *out_offsets, int64_t *out_index, void *out_value, int64_t *total,
                                int32_t *status, void *ws, size_t ws_bytes, void *stream);

/* out_value (optional, content dtype, n items): the extreme value of every event that has one, in the same dense
 * order as out_index -- i.e. the "leading object" a[a.argmax()] (jagged.py:541-560 applied to the result) without
 * the gather pass.                                                                                                 */
/* turn best[] into the dense result: out_offsets[n+1] (0/1 counts scanned) and out_index[<=n]; *total = count.   */
int jg_compact_nonneg(const int64_t *best, int64_t n, int64_t *out_offsets, int64_t *out_index, int64_t *total,
                      void *ws, size_t ws_bytes, void *stream);

/* ---- jagged __getitem__ ---------------------------------------------------------------------------------------- */
/* jagged.py:562-589  a[jagged_bool]: single pass (per-event popcount -> lookback scan -> compaction).
 * mask and content are both addressed through `offsets` (absolute positions offsets[0]..offsets[n]);
 * mask_shift = mask_offsets[0] - offsets[0] re-bases a mask whose content starts elsewhere.
 * out must hold offsets[n]-offsets[0] items (upper bound); *total receives K.                                     */
int jg_mask_select(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift,
                   const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                   int64_t *new_counts /* n entries, may be NULL: the result's counts for free */, int64_t *total,
                   void *ws, size_t ws_bytes, void *stream);
/* The same selection in two calls: phase 1 queues the counting pass and the scan of the tile totals (after it *total
 * holds K), phase 2 the compaction (same arguments and workspace); phase 0 is jg_mask_select.  (phase 3: the whole
 * selection as ONE pass over the data on published tile aggregates -- same result, measured slower on B200, kept as
 * the documented alternative: profiles/r2_lag_study.md section 9.)  The host reads K on a
 * side stream as soon as phase 1 has run and allocates / returns the result while the compaction is still running, so
 * the operation that follows is queued behind it without the device going idle (jagged.py:562-589 returns only after
 * numpy has finished; on a device the data-dependent size is the only thing the host ever waits for).              */
int jg_mask_select_phase(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift,
                         const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                         int64_t *total, void *ws, size_t ws_bytes, int phase, void *stream);
/* counts of two layouts must agree (jagged.py:911-912): raises JG_ST_COUNTS_MISMATCH                               */
int jg_check_same_counts(const int64_t *offsets_a, const int64_t *offsets_b, int64_t n, int32_t *status,
                         void *stream);
/* jagged.py:541-560  a[jagged_int]: out[k] = content[offsets[p] + wrap(index[k])], p = event of k in
 * index_offsets; negative wrap (:552-553); JG_ST_INDEX_OOB (:555-556).  index_dtype in {JG_I32, JG_I64}.          */
int jg_index_gather(const void *content, int itemsize, const int64_t *offsets, const void *index,
                    int index_dtype, const int64_t *index_offsets, int64_t n, void *out, int32_t *status,
                    void *stream);
/* flat boolean selection content[mask] (event-level selection a[flat_bool], jagged.py:599-605): positions
 * (caller scratch, int64[n+1]) receives the exclusive scan of the mask, *total the number selected; content may
 * be NULL to run the scan only.  out must hold n items (upper bound).                                              */
int jg_flat_compact(const void *content, int itemsize, const uint8_t *mask, int64_t n, void *out, int64_t *positions,
                    int64_t *total, void *ws, size_t ws_bytes, void *stream);
/* a[a OP c] in one call (jagged.py:944-1051 comparison ufunc + jagged.py:562-589 selection): the comparison of the
 * float column `key` (JG_F32 / JG_F64, positioned like `content` shifted by key_shift items) with a scalar is
 * evaluated inside the counting and the compaction sweeps, so the M-byte mask is neither written nor read.  op is
 * JG_OP_GT / GE / LT / LE with the column on the left (the host mirrors `c OP a`).  When key == content the column is
 * staged once; a content column other than the key must have 4- or 8-byte items.  The rest is jg_mask_select.     */
int jg_compare_select(const void *content, int itemsize, const void *key, int key_dtype, int op, double scalar,
                      int64_t key_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                      int64_t *new_counts, int64_t *total, void *ws, size_t ws_bytes, void *stream);

/* jg_compare_select in two calls (see jg_mask_select_phase).  phase 4 (float32 column compared with itself; like phase 0
 * otherwise): the whole selection as ONE launch in which a CTA counts a range of 8 event tiles from global memory, takes
 * the range's base from the published aggregates and then compacts its tiles from L2 -- the fifth single-pass form,
 * bit-exact, measured slower than the three launches (1.96 vs 1.58 ms) and kept as that measured alternative.          */
int jg_compare_select_phase(const void *content, int itemsize, const void *key, int key_dtype, int op, double scalar,
                            int64_t key_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                            int64_t *new_counts, int64_t *total, void *ws, size_t ws_bytes, int phase, void *stream);

/* scatter content[index[k]] = values[k], or the `itemsize` low bytes of scalar_bits when values is NULL: the store
 * of __setitem__ (jagged.py:789-838: `self._content[indexes] = what`) and of parents2startsstops (jagged.py:73-93:
 * `starts[where[real]] = ...`, for which skip_negative != 0 drops the negative parents).  No bounds check: the
 * indexes come from jg_index_gather / jg_mask_select on absolute positions, which are checked there.              */
int jg_put(void *content, int itemsize, const int64_t *index, int64_t n, const void *values, uint64_t scalar_bits,
           int skip_negative, void *stream);
/* plain gather out[k] = content[index[k]] (jagged.py:1294 content[left]); no bounds check beyond content_len      */
int jg_take(const void *content, int itemsize, const int64_t *index, int64_t n_index, void *out, void *stream);

/* ---- element-tiled variants for long events ------------------------------------------------------------------------
 * jagged.py:1552-1563 (the per-event reduceat loop) and its siblings when the COUNT DISTRIBUTION has long events:
 * the host calls these instead of the event-tiled entry points above when content_len > 7 * n (north_star: "variants
 * chosen by count distribution").  Work is split into tiles of 16 KB of content instead of 512 events, so Poisson(40)
 * counts, heavy tails and a single event of 2^31 elements all keep every SM busy; an event that spans several tiles
 * is merged by a second, tiny launch.  Results are identical to the event-tiled calls (float sum / prod within the
 * reduction-order tolerance).  `content_len` = offsets[n] - offsets[0]; `ws` holds jg_elem_workspace_bytes bytes.     */
size_t jg_elem_workspace_bytes(int64_t content_len, int itemsize);
/* jagged.py:1459-1565: as jg_segreduce; dtype in {JG_F32, JG_F64, JG_I32, JG_I64}                                      */
int jg_segreduce_elem(const void *content, int dtype, const int64_t *offsets, int64_t n, int op, void *out,
                      int64_t content_len, void *ws, size_t ws_bytes, void *stream);
/* jagged.py:1567-1601: as jg_segargminmax (best[] and best_value[] both required); jg_compact_nonneg makes it dense  */
int jg_segargminmax_elem(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min, int64_t *best,
                         void *best_value, int64_t content_len, void *ws, size_t ws_bytes, void *stream);
/* jagged.py:562-589: as jg_mask_select / jg_compare_select for 4- and 8-byte items: ONE persistent launch (flat stream
 * compaction on lagged tile prefixes; an event reads its new offset from the on-chip prefix of the tile it starts in) */
int jg_mask_select_elem(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift, const int64_t *offsets,
                        int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts, int64_t *total,
                        int64_t content_len, void *ws, size_t ws_bytes, void *stream);
int jg_compare_select_elem(const void *content, int itemsize, const void *key, int key_dtype, int op, double scalar,
                           int64_t key_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                           int64_t *new_counts, int64_t *total, int64_t content_len, void *ws, size_t ws_bytes,
                           void *stream);
/* jagged.py:49-54, 430-443, 840-881: mode 0 = offsets2parents (int64 out, absolute positions, -1 below offsets[0]),
 * 1 = localindex (int64 out), 2 = broadcast of flat[n] (`itemsize` bytes per item) onto the content positions         */
int jg_expand_elem(int mode, const void *flat, int itemsize, const int64_t *offsets, int64_t n, void *out,
                   int64_t content_len, void *ws, size_t ws_bytes, void *stream);

/* ---- per-event argsort ------------------------------------------------------------------ */
/* jagged.py:1603-1631  out[offsets[i]-offsets[0] + r] = local index of the element of event i with rank r: stable
 * by value (descending unless `ascending`), ties and -0.0/0.0 in position order, NaN last in both directions.
 * Events longer than 2048 elements are not sorted here: JG_ST_SORT_LONG is raised in *status, long_list[0] counts
 * them and long_list[1 .. long_cap] holds their event indices (long_list[0] must be zero on entry; a caller that
 * sizes long_cap = (offsets[n]-offsets[0]) / 2049 + 1 can never overflow it); sort each with jg_argsort_long.       */
int jg_argsort(const void *content, int dtype, const int64_t *offsets, int64_t n, int ascending, int64_t *out,
               int32_t *status, int64_t *long_list, int64_t long_cap, void *stream);
/* one long event: content[start .. start+count) -> out[0 .. count) (local indices); ws >= jg_argsort_long_workspace  */
int64_t jg_argsort_long_workspace(int64_t count);
int jg_argsort_long(const void *content, int dtype, int64_t start, int64_t count, int ascending, int64_t *out,
                    void *ws, size_t ws_bytes, void *stream);

/* ---- pad / fillna ----------------------------------------------------------------------- */
/* jagged.py:1851-1900  pad(): event i of the padded (dense) layout dst_offsets holds the first counts[i] items of
 * content[starts[i] ...] followed by zeros; invalid[j] = 1 on the padding (optional).  The caller chooses
 * dst_offsets: scan of max(counts, length) for clip=False, i*length with counts clipped to length for clip=True.   */
int jg_pad(const void *content, int itemsize, const int64_t *starts, const int64_t *counts, const int64_t *dst_offsets,
           int64_t n, void *out, uint8_t *invalid, void *stream);
/* masked.py:401-406 + base.py:645-654  fillna(): out = value where (mask != 0) == masked_when or content is NaN,
 * else content; mask may be NULL (NaN replacement only).  The value arrives as double and int64 like jg_binary's.  */
int jg_fill_masked(const void *content, int dtype, const uint8_t *mask, int masked_when, double value_f64,
                   int64_t value_i64, void *out, int64_t len, void *stream);

/* ---- Arrow ingest ----------------------------------------------------------------------- */
/* arrow.py:380-395  boolean values arrive bit-packed, LSB first: out[i] = bit (bit_offset + i) of `bits` as 0/1
 * bytes; `bits` must hold ceil((bit_offset + n) / 8) bytes.  List offsets (int32 / int64) need no kernel of their
 * own: JaggedArray.fromoffsets takes them as they are (jagged.py:142-153).                                          */
int jg_unpack_bits(const uint8_t *bits, int64_t bit_offset, int64_t n, uint8_t *out, void *stream);

/* ---- concatenation ---------------------------------------------------------------------- */
/* jagged.py:1633-1649  concatenate(axis=0): out[i] = offsets[i] + delta for `count` entries (offsets of a later
 * array shifted by the number of content items in front of it)                                                      */
int jg_offsets_rebase(const int64_t *offsets, int64_t count, int64_t delta, int64_t *out, void *stream);
/* jagged.py:1651-1733  concatenate(axis=1): out[dst_starts[i] + k] = content[offsets[i] + k] -- event i of a dense
 * input lands at its slot inside the merged event                                                                    */
int jg_segment_scatter(const void *content, int itemsize, const int64_t *offsets, const int64_t *dst_starts, int64_t n,
                       void *out, void *stream);
/* jagged.py:1651-1733  concatenate(axis=1) of k = 2..4 dense inputs in one sweep over the output layout:
 * out[out_offsets[i] + (items of inputs < q in event i) + t] = contents[q][offsets[q][i] + t].  `contents` and `offsets`
 * are HOST arrays of k device pointers; every content has `itemsize`; out_offsets = scan of the summed counts.           */
int jg_segment_merge(const void *const *contents, const int64_t *const *offsets, int k, int itemsize,
                     const int64_t *out_offsets, int64_t n, void *out, void *stream);

/* ---- broadcasting and elementwise ufuncs ------------------------------------------------------------------------ */
/* jagged.py:866-875  tojagged(flat): out[j - offsets[0]] = flat[parent(j)]                                          */
int jg_broadcast(const void *flat, int itemsize, const int64_t *offsets, int64_t n, void *out, void *stream);
/* jagged.py:1043  ufunc(lhs, rhs) on dense contents.  Operands are converted to `compute_dtype` on load, the op
 * runs in that type, and the result is stored as `out_dtype` (comparisons -> JG_B8).  rhs_kind selects array /
 * scalar / per-event broadcast (then `offsets`/`n` describe the events of lhs and lhs starts at offsets[0]).
 * swap != 0 computes op(rhs, lhs) (reflected operators).                                                             */
int jg_binary(int op, const void *lhs, int lhs_dtype, const void *rhs, int rhs_dtype, int rhs_kind,
              double scalar_f64, int64_t scalar_i64, int swap, int compute_dtype, void *out, int out_dtype,
              int64_t len, const int64_t *offsets, int64_t n, void *stream);
int jg_unary(int op, const void *in, int in_dtype, int compute_dtype, void *out, int out_dtype, int64_t len,
             void *stream);
/* whole-array reduction of a flat device array to one value (for global totals before an NCCL all-reduce)          */
int jg_flat_reduce(const void *in, int dtype, int64_t len, int op, void *out /* device, out dtype as segreduce */,
                   void *ws, size_t ws_bytes, void *stream);

/* ---- combinatorics ------------------------------------------------------------------------------------------------ */
/* single pass: per-event output count -> lookback scan -> index generation.
 *   kind = JG_COMB_PAIRS / DISTINCTS / CHOOSE(k) / CROSS (offsets_b only for CROSS).
 * Phase 1 (jg_comb_offsets): out_offsets[n+1] and *total = P (device).  Phase 2 (jg_comb_fill): writes k local
 * index columns cols[c][P] (int64), or, if `absolute` != 0, indices with the event start added
 * (jagged.py:1086-1087, 1319-1320).                                                                                   */
int jg_comb_offsets(int kind, int k, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                    int64_t *out_offsets, int64_t *total, void *ws, size_t ws_bytes, void *stream);
int jg_comb_fill(int kind, int k, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                 const int64_t *out_offsets, int64_t *const *cols /* host array of k device pointers */,
                 int absolute, void *stream);
/* fused value form (jagged.py:1294,1350): out_l[p] = content_a[start_a + i], out_r[p] = content_b[start_b + j]
 * for kind in PAIRS / DISTINCTS / CROSS without materialising the index columns.                                     */
int jg_comb_gather2(int kind, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                    const int64_t *out_offsets, const void *content_a, int itemsize_a, const void *content_b,
                    int itemsize_b, void *out_l, void *out_r, void *stream);

/* record form of the above (Table content: jagged.py:1294/1350 with the row views of table.py:619-629): up to 4
 * leaf columns per side, all of one itemsize, gathered in one sweep (host arrays of device pointers).
 * For PAIRS / DISTINCTS pass the same input columns on both sides.                                                  */
int jg_comb_gather_multi(int kind, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                         const int64_t *out_offsets, int itemsize, int n_left, const void *const *in_left,
                         void *const *out_left, int n_right, const void *const *in_right, void *const *out_right,
                         void *stream);

/* Fused value form + ufunc chain ("pair mass fused from index generation"): replaces, for ONE result
 * column, the sequence  pairs()/distincts()/cross() (jagged.py:1232-1367) -> column access (table.py:587-600) ->
 * a chain of numpy ufuncs through JaggedArray.__array_ufunc__ (jagged.py:944-1051) that the reference runs as one
 * numpy pass (and one P-sized temporary) per step.  The expression is an accumulator program, one 32-bit word per
 * step:  op | operand kind << 8 | reverse << 12 | operand index << 16
 *   op = a binary / comparison JG_OP_* code: acc = op(acc, operand)  (reverse: op(operand, acc))
 *      = a unary JG_OP_* code             : acc = op(acc)
 *      = JG_EV_LOAD : acc = operand        JG_EV_STORE : slot[index] = acc
 * All columns and the result have the compute dtype (JG_F32 or JG_F64); with out_bool the final accumulator
 * (0 / 1 from a comparison) is stored as one byte.  The first step must be a JG_EV_LOAD.                          */
enum {
    JG_EV_LOAD = 30,         /* (unused JG_OP_* codes below the unary range) */
    JG_EV_STORE = 31
};
enum {
    JG_EV_SLOT = 0,          /* operand = spill slot[index]                                               */
    JG_EV_LEFT = 1,          /* operand = left[index][position of the combination's first item]          */
    JG_EV_RIGHT = 2,         /* operand = right[index][position of the combination's second item]        */
    JG_EV_POS = 3,           /* operand = pos[index][output position] (an already materialised P-sized column) */
    JG_EV_CONST = 4          /* operand = consts[index]                                                   */
};
#define JG_EVAL_MAX_INS 48
#define JG_EVAL_MAX_COLS 8
#define JG_EVAL_MAX_CONSTS 8
#define JG_EVAL_MAX_SLOTS 10
typedef struct JgEvalProgram {
    uint32_t ins[JG_EVAL_MAX_INS];
    const void *left[JG_EVAL_MAX_COLS];   /* device pointers, indexed like the left collection's content  */
    const void *right[JG_EVAL_MAX_COLS];  /* ... the right collection's content (PAIRS / DISTINCTS: the same collection) */
    const void *pos[JG_EVAL_MAX_COLS];    /* device pointers, indexed by output position */
    double consts[JG_EVAL_MAX_CONSTS];
    int32_t nins, nslots, out_bool, reserved;
} JgEvalProgram;
int jg_comb_eval(int kind, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n, const int64_t *out_offsets,
                 int dtype, const JgEvalProgram *program /* host */, void *out, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* JAGGED_B200_H */
