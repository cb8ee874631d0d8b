// shift_table.cuh -- the FP32 shift-table kernel of the table path (see table_path.cu for the
// algebra) and its per-(S, Nw) launchers.  The kernel is instantiated for S = 3..19 and one window
// half-width per translation unit: shift_table_inst.cu is compiled once per UMPA_INST_NW value
// (-1 = plain table, 0..6 = filtered) so that the instantiations build in parallel.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

#ifndef UMPA_FFMA2_MAX_S
#define UMPA_FFMA2_MAX_S 99      // packed FFMA2 in the main loop for S <= this (experiments: 0 = plain FFMA everywhere)
#endif
#ifndef UMPA_COLPASS4
#define UMPA_COLPASS4 1          // 1: column pass of the filter epilogue on four output rows per thread (see the kernel)
#endif

namespace shift_table {

constexpr int SMEM_CAP = 227 * 1024;

// ------------------------------------------------------------------ shift tables
//
// table[s][p] = sum_k A_k(p+s) * B_k(p)   (FILTER: then window-filtered over p)
//
// A persistent CTA works through a list of ITEMS; an item is a column strip (TW outputs wide) times a segment
// of output rows, processed top to bottom in CHUNKS of EH rows.  Per chunk the CTA produces ALL S*S shifts:
//   * the chunk (EH rows x 32 columns = TW + 2*halo) is cut into strips of 4 consecutive pixels; a thread owns
//     one strip -> 8 strips per row, so every quarter warp reads one contiguous 128 B shared-memory line
//     (conflict-free LDS.128);
//   * G warp groups work on the same frame at the same time, group g accumulating shift rows
//     [ (pass*G+g)*SH, +SH ): SH*S*4 FP32 accumulators per thread live in registers across all
//     frames, so each frame tile is streamed through shared memory exactly once per pass;
//   * the FMAs are Blackwell's packed FFMA2 (fma.rn.f32x2): the products A(q) * B(x) and A(q) * B(x+1) of one
//     reference value with two neighbouring sample pixels belong to the shifts q - x and q - x - 1, so one
//     instruction with A(q) as the broadcast scalar operand and the pair (B(x), B(x+1)) -- a natural register
//     pair of the LDS.128 -- does both: 20 issue slots per shift row instead of 36.  The FMA pipe does the same
//     work either way (measured, tools/probe/ffma2_probe.cu); what it buys is issue slots next to the ten LDS.128
//     per frame: 605 -> 505 cycles per frame and SM, against 480 from shared-memory bandwidth;
//   * frames arrive by TMA (cp.async.bulk.tensor, 3-D map over [Na][H][pitch], zero fill out
//     of bounds) into a ring of NST stages of FB frames each (one box per stack and stage),
//     guarded by full/empty mbarriers; thread 0 is the producer, nobody executes a per-frame
//     __syncthreads();
//   * epilogue (per shift row, window half-width compiled in): row pass of the separable window in
//     registers (shuffles inside the 8-lane row group), row-filtered strips to the group's slice of
//     shared memory, named barrier per group, column pass per thread (K float4 loads), float4
//     stores to the table.
//   * STREAMING: the column pass of an output row needs the row-filtered chunk rows r .. r+2*halo.  The last
//     2*halo row-filtered rows of a chunk are kept in shared memory (`carry`, all S*S planes) for the next chunk
//     of the same item, so a chunk of EH rows yields EH output rows: the window halo in y is paid once per
//     segment instead of once per tile (config 2: 12 of 16 rows -> 16 of 16).  Where the carry does not fit
//     (large S * Nw) the host makes every segment one chunk high (EH - 2*halo rows): the classic halo tile.

struct TableParams {
    float *table;                // entry (row, shift, col) at table[row * row_stride + shift * plane_stride + col]:
                                 // the S*S shifts of one pixel row lie next to each other (and the two tables of a DF
                                 // model next to each other), so a walk touches ONE region of a few hundred KB
    size_t row_stride;           // floats between consecutive pixel rows
    int plane_stride;            // floats between consecutive shifts of one row (the common row pitch of the tables)
    const float *g;              // window factor (FILTER only)
    int Na, Nw;
    int oy, ox;                  // raw coordinates of table element (0,0)
    int rows, cols_p;            // table rows (exact), padded columns (nstrips * TW)
    int TW;                      // output columns per strip
    int EH;                      // chunk rows (chunk columns are EXT_W)
    int seg_rows, nseg, nstrips; // items: nstrips x nseg segments of seg_rows output rows (the last one may be shorter)
    int nchunk_full, nchunk_last;// chunks of a full segment / of the last one (host: no integer division per item)
    int stream;                  // (host) 0 halo tiles, 1 streaming chunk-major, 2 streaming pass-major
    int pass_major;              // streaming when all S*S carry planes do not fit: the passes of an item become the outer
                                 // loop (pass -> chunk instead of chunk -> pass), so that only the G*SH*S planes of the
                                 // current pass need a carry (config 4: 60 of 225 planes, 46 KB)
    int AH, AP;                  // A tile rows, pitch (= TMA box width)
    int G, npass, nstage;
    int a_stage_floats, stage_floats;   // per-stage layout: A tile then B tile (128 B aligned)
    int FB;                             // frames per TMA box / ring stage
    int dbg;                            // experiments only (UMPA_TAB_DBG): 1 = skip the epilogue, 2 = skip the FMA loop
};

constexpr int EXT_W = 32;          // chunk width: one 128 B line per row
constexpr int MAX_NT = 384;        // threads per CTA (3 groups x 16 rows x 8 strips)
constexpr int MAX_STAGES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tiled TMA load: box (c0.., c1.., c2) of the tensor map -> dense smem tile, completes on `bar`
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
            "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// NWT: -1 = plain table (no window filter); >= 0 = window half-width, filter fully unrolled
template <int S, int SH, int NWT>
__global__ void __launch_bounds__(MAX_NT)
shift_table_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TableParams p)
{
    extern __shared__ __align__(128) float sm[];
    constexpr int HS = (S - 1) / 2;                  // max |shift|
    // TMA needs the innermost box coordinate 16 B aligned (measured: unaligned -> illegal
    // instruction).  The B tile origin is aligned by construction (host shifts the table
    // origin); the A tile starts HS + DELTA columns to its left so that it is aligned too.
    constexpr int DELTA = (4 - HS % 4) % 4;
    constexpr int NA4 = (DELTA + S + 3 + 3) / 4;     // float4 loads covering DELTA+S+3 floats of an A row
    constexpr bool FILTER = NWT >= 0;
    constexpr int K = FILTER ? 2 * NWT + 1 : 1;
    constexpr int HALO = FILTER ? NWT : 0, H2 = 2 * HALO;
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];

    const int tid = threadIdx.x, nt = blockDim.x;
    const int TG = p.EH * (EXT_W / 4);               // threads per group
    const int grp = tid / TG, lt = tid - grp * TG;
    const int er = lt >> 3, ec = (lt & 7) << 2;      // strip: chunk row, first chunk column
    float *cbuf = sm + (size_t)p.nstage * p.stage_floats;                // [G*S][EH][EXT_W] (FILTER only)
    float *carry = cbuf + (size_t)p.G * S * p.EH * EXT_W;                // [S*S][H2][EXT_W] (FILTER, segments of > 1 chunk)
    const uint32_t stage_bytes = (uint32_t)(p.FB * (p.AH * p.AP + p.EH * EXT_W)) * sizeof(float);
    const int a_frame = p.AH * p.AP, b_frame = p.EH * EXT_W;             // one frame inside a stage

    // BIG: 120 accumulators per thread (S >= 11).  Registers that are live across the frame loop then cost the A rows in
    // flight (see the producer state below), so the epilogue's per-chunk quantities and the window factors are
    // derived after the frame loop; with 108 accumulators (S <= 9) deriving them early is what ptxas schedules best
    // (measured both ways on both: config 2's cross table 0.87 vs 0.99 ms, config 4's 29.4 vs 24.3 ms).
    constexpr bool BIG = SH * S * 4 > 112;
    // pass-major order (TableParams::pass_major) exists where a table can need more than one pass at the streaming
    // chunk height: S >= 11.  Compiled out below that -- the order selects cost the S = 9 kernel 10 % (config 2).
    constexpr bool PMC = S >= 11;
    const bool pass_major = PMC && p.pass_major;
    float gk_early[K];
#pragma unroll
    for (int v = 0; v < K; v++) gk_early[v] = (FILTER && !BIG) ? __ldg(p.g + v) : 1.f;
    if (tid == 0) {
        for (int s = 0; s < p.nstage; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], nt / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    // ---- persistent CTA: items blockIdx.x, +gridDim.x, ... ; the frame ring runs across chunks and items ----
    // item -> (strip, segment); neighbouring CTAs work on neighbouring strips at the same time (shared x halos in L2)
    // (segment, strip) of an item advance by gridDim.x items at a time: carried along instead of divided out
    const int nitems = p.nstrips * p.nseg;
    const int step_seg = (int)gridDim.x / p.nstrips, step_strip = (int)gridDim.x - step_seg * p.nstrips;
    auto advance = [&](int &seg, int &strip) {
        seg += step_seg; strip += step_strip;
        if (strip >= p.nstrips) { strip -= p.nstrips; seg++; }
    };
    auto chunks_of_seg = [&](int seg) { return seg == p.nseg - 1 ? p.nchunk_last : p.nchunk_full; };
    const int seg_first = (int)blockIdx.x / p.nstrips, strip_first = (int)blockIdx.x - seg_first * p.nstrips;
    // A ring stage holds FB consecutive frames (one TMA box per stack): a box costs ~450-500 cycles of
    // TMA-unit time whatever its size (measured: A only, B only and both take the same time with the FMA
    // loop and the epilogue switched off), so one box per frame capped the kernel at ~515 cycles per frame.
    const int nbox = (p.Na + p.FB - 1) / p.FB;       // boxes per pass over the frames (the last one may run past Na: zero fill)
    const int per_chunk = p.npass * nbox;

    // producer state (thread 0): next (item, chunk, frame) to request, and where.  Only one thread uses it, once per
    // box, but as local variables it holds a dozen registers in EVERY thread across the frame loop.  With 120
    // accumulators per thread (S >= 11) those registers are what the A rows in flight need: with the state in registers
    // every A row load waited for the previous row's FMAs and config 4 (S = 15) lost 20 %, so there it lives in shared
    // memory.  With 108 accumulators (S <= 9) there is room, and the serial shared-memory round trips of thread 0 would
    // make warp 0 the straggler of every barrier (config 2: +6 %): registers.
    struct Producer { int total, issued, item, chunks, left, frame, stage, ax, ay, bx, by, seg, strip, inner, outer; };
    __shared__ Producer pr_shared;
    Producer pr_local;
    Producer &pr = BIG ? pr_shared : pr_local;
    auto pr_coords = [&]() {                         // first chunk of pr.item = (pr.seg, pr.strip)
        if (pr.item >= nitems) return;
        pr.by = p.oy + pr.seg * p.seg_rows - HALO; pr.bx = p.ox + pr.strip * p.TW - HALO;
        pr.ay = pr.by - HS; pr.ax = pr.bx - HS - DELTA;
        pr.chunks = chunks_of_seg(pr.seg);
    };
    auto issue_next = [&]() {
        const int st = pr.stage;
        float *As = sm + (size_t)st * p.stage_floats, *Bs = As + p.a_stage_floats;
        mbar_expect_tx(&full_bar[st], stage_bytes);
        tma_load_3d(As, &mapA, pr.ax, pr.ay, pr.frame, &full_bar[st]);
        tma_load_3d(Bs, &mapB, pr.bx, pr.by, pr.frame, &full_bar[st]);
        pr.issued++;
        pr.stage = st + 1 == p.nstage ? 0 : st + 1;
        pr.frame = pr.frame + p.FB >= p.Na ? 0 : pr.frame + p.FB;
        if (--pr.left == 0) {                        // next (chunk, pass) of the item in the consumers' order, or the next item
            if constexpr (!PMC) {                    // chunk -> pass, counted as one run of boxes per chunk
                pr.left = per_chunk;
                if (--pr.chunks > 0) { pr.by += p.EH; pr.ay += p.EH; }
                else { pr.item += gridDim.x; advance(pr.seg, pr.strip); pr_coords(); }
            } else {
                pr.left = nbox;
                bool item_done = false;
                if (!pass_major) {                   // chunk -> pass: the tile moves down after the last pass of a chunk
                    if (++pr.inner == p.npass) { pr.inner = 0; pr.by += p.EH; pr.ay += p.EH; item_done = ++pr.outer == pr.chunks; }
                } else {                             // pass -> chunk: down after every unit, back up after the last chunk
                    pr.by += p.EH; pr.ay += p.EH;
                    if (++pr.inner == pr.chunks) {
                        pr.inner = 0; pr.by -= pr.chunks * p.EH; pr.ay -= pr.chunks * p.EH;
                        item_done = ++pr.outer == p.npass;
                    }
                }
                if (item_done) { pr.outer = 0; pr.item += gridDim.x; advance(pr.seg, pr.strip); pr_coords(); }
            }
        }
    };
    if (tid == 0) {
        int total = 0;
        for (int it = blockIdx.x, sg = seg_first, sp = strip_first; it < nitems; it += gridDim.x, advance(sg, sp))
            total += chunks_of_seg(sg) * per_chunk;
        pr.total = total; pr.issued = 0; pr.item = blockIdx.x; pr.left = PMC ? nbox : per_chunk; pr.frame = 0; pr.stage = 0;
        pr.seg = seg_first; pr.strip = strip_first; pr.chunks = 0; pr.inner = 0; pr.outer = 0;
        pr_coords();
        for (int n = 0; n < p.nstage && n < total; n++) issue_next();
    }

    // What the epilogue needs to know about a chunk.  Output row of this thread: the column pass ends at its own chunk
    // row (taps er-H2 .. er); the four-row column pass works on groups of four chunk rows e4 .. e4+3 (output rows
    // r_out4 .. r_out4+3, row_ok: which of them are stored).
    struct Epi { int r_out, r_out4; unsigned row_ok; bool store; };
    auto make_epi = [&](int chunk, int item_rows) {
        Epi e;
        e.r_out = chunk * p.EH + er - H2;
        e.store = e.r_out >= 0 && e.r_out < item_rows && ec < p.TW;
        e.r_out4 = chunk * p.EH + (er & ~3) - H2;
        e.row_ok = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) e.row_ok |= (e.r_out4 + i >= 0 && e.r_out4 + i < item_rows && ec < p.TW) ? 1u << i : 0u;
        return e;
    };

    int stage = 0, phase = 0;                        // consumer ring position
    int prev_stage = 0, prev_phase = 0;
    bool first = true;
    // Accumulators of one thread: strip pixels x = 0..3, shift rows sh < SH, shift columns sj < S.  For the pixel
    // pair (x0, x0+1), x0 = 2 xp, the FFMA2 pair t (1 <= t < S) holds (shift t of pixel x0, shift t-1 of pixel
    // x0+1): both are fed by reference element DELTA + t + x0.  Shift 0 of x0 and shift S-1 of x0+1 have no
    // partner (plain FFMA).  ACC(sh, sj, x) names the register of one (shift, pixel).
    float2 accp[SH][2][S - 1];
    float accs[SH][2][2];
#define ACC(SH_, SJ_, PX_) (((PX_) & 1) ? ((SJ_) == S - 1 ? accs[SH_][(PX_) >> 1][1] : accp[SH_][(PX_) >> 1][(SJ_)].y) \
                                        : ((SJ_) == 0 ? accs[SH_][(PX_) >> 1][0] : accp[SH_][(PX_) >> 1][(SJ_) - 1].x))

    for (int item = blockIdx.x, seg = seg_first, strip = strip_first; item < nitems; item += gridDim.x, advance(seg, strip)) {
      const int row0 = seg * p.seg_rows, tx0 = strip * p.TW;             // table coords of the item
      const int item_rows = min(p.seg_rows, p.rows - row0), nchunk = chunks_of_seg(seg);
      const int n_outer = pass_major ? p.npass : nchunk, n_inner = pass_major ? nchunk : p.npass;
      for (int outer = 0; outer < n_outer; outer++) {
        Epi epi_early{};
        if (!PMC) epi_early = make_epi(outer, item_rows);
        for (int inner = 0; inner < n_inner; inner++) {
            const int chunk = pass_major ? inner : outer, pass = pass_major ? outer : inner;
            if (PMC && !BIG) epi_early = make_epi(chunk, item_rows);
            const int si0 = (pass * p.G + grp) * SH; // first shift row of this thread in this pass
            const bool work = si0 < S;
#pragma unroll
            for (int a = 0; a < SH; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
#pragma unroll
                    for (int c = 0; c < S - 1; c++) accp[a][b][c] = make_float2(0.f, 0.f);
                    accs[a][b][0] = accs[a][b][1] = 0.f;
                }

            for (int box = 0; box < nbox; box++) {
                if (tid == 0 && !first && pr.issued < pr.total) {        // refill the stage the previous box used
                    mbar_wait(&empty_bar[prev_stage], prev_phase);
                    issue_next();
                }
                first = false;
                mbar_wait(&full_bar[stage], phase);
                const int nfr = min(p.FB, p.Na - box * p.FB);
                if (work && !(p.dbg & 2))
                  for (int fr = 0; fr < nfr; fr++) {
                    const float *As = sm + (size_t)stage * p.stage_floats + fr * a_frame;
                    const float *Bs = sm + (size_t)stage * p.stage_floats + p.a_stage_floats + fr * b_frame;
                    const float4 b4 = *reinterpret_cast<const float4 *>(Bs + er * EXT_W + ec);
                    const float2 bp[2] = {make_float2(b4.x, b4.y), make_float2(b4.z, b4.w)};
                    const float *arow0 = As + (er + si0) * p.AP + ec;
#pragma unroll
                    for (int sh = 0; sh < SH; sh++) {
                        if (si0 + sh < S) {
                            const float *arow = arow0 + sh * p.AP;
                            float av[4 * NA4];
#pragma unroll
                            for (int v = 0; v < NA4; v++) {
                                const float4 t = *reinterpret_cast<const float4 *>(arow + 4 * v);
                                av[4 * v] = t.x; av[4 * v + 1] = t.y; av[4 * v + 2] = t.z; av[4 * v + 3] = t.w;
                            }
#pragma unroll
                            for (int xp = 0; xp < 2; xp++) {
                                accs[sh][xp][0] = fmaf(bp[xp].x, av[DELTA + 2 * xp], accs[sh][xp][0]);
#pragma unroll
                                for (int t = 1; t < S; t++) {
                                    const float a = av[DELTA + t + 2 * xp];
                                    if (S <= UMPA_FFMA2_MAX_S)
                                        accp[sh][xp][t - 1] = __ffma2_rn(make_float2(a, a), bp[xp], accp[sh][xp][t - 1]);
                                    else {
                                        accp[sh][xp][t - 1].x = fmaf(a, bp[xp].x, accp[sh][xp][t - 1].x);
                                        accp[sh][xp][t - 1].y = fmaf(a, bp[xp].y, accp[sh][xp][t - 1].y);
                                    }
                                }
                                accs[sh][xp][1] = fmaf(bp[xp].y, av[DELTA + S + 2 * xp], accs[sh][xp][1]);
                            }
                        }
                    }
                  }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);     // this warp is done with the stage
                prev_stage = stage; prev_phase = phase;
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
            }

            // ---------------- epilogue of this pass ----------------
            const Epi epi = BIG ? make_epi(chunk, item_rows) : epi_early;
            const int r_out = epi.r_out;
            const bool store = epi.store;
#if UMPA_COLPASS4
            const int e4 = er & ~3, qi = er & 3, r_out4 = epi.r_out4;
            const unsigned row_ok = epi.row_ok;
            const bool any_store = row_ok != 0;
#endif
            float gk[K];
#pragma unroll
            for (int v = 0; v < K; v++) gk[v] = (FILTER && BIG) ? __ldg(p.g + v) : gk_early[v];
            if (p.dbg & 1) {
                if (work && tid == 0x7fffffff) p.table[0] = ACC(0, 0, 0);      // keeps the accumulators alive
            } else if (!FILTER) {
                if (work && store) {
#pragma unroll
                    for (int sh = 0; sh < SH; sh++) {
                        const int si = si0 + sh;
                        if (si < S) {
                            float *dst = p.table + (size_t)(row0 + r_out) * p.row_stride + (size_t)(si * S) * p.plane_stride + tx0 + ec;
#pragma unroll
                            for (int sj = 0; sj < S; sj++)
                                *reinterpret_cast<float4 *>(dst + sj * p.plane_stride) =
                                    make_float4(ACC(sh, sj, 0), ACC(sh, sj, 1), ACC(sh, sj, 2), ACC(sh, sj, 3));
                        }
                    }
                }
            } else {
                // Separable window filter, one shift row (S planes per group) at a time.
                //  row pass: in registers; the 8 lanes of a quarter warp hold one chunk row, output x
                //            needs columns x .. x+2Nw = own strip + the next NSH strips (shuffles; lanes past
                //            the row end only feed outputs x >= TW, which are never stored);
                //  column pass: through the group's slice of cbuf (conflict-free float4 lines): each thread filters
                //            its own strip over the chunk rows er-2Nw .. er and stores one float4 of the table; the
                //            first 2Nw rows of a chunk reach back into the previous chunk's rows (carry);
                //  carry:    the threads of the last 2Nw rows then copy their row-filtered strips to the carry.
                // Only the 128 threads of a group share data: named barrier per group, no __syncthreads.
                constexpr int NSH = (K - 1 + 3) / 4;
                const int plane = p.EH * EXT_W;
                float *cg = cbuf + (size_t)grp * S * plane + er * EXT_W + ec;
                const bool keep = chunk + 1 < nchunk && er >= p.EH - H2;
#pragma unroll
                for (int sh = 0; sh < SH; sh++) {
                    const int si = si0 + sh;
                    if (!work || si >= S) continue;              // uniform per group
                    // first carry plane of this shift row: all S*S planes are kept, or only those of the current pass
                    const int cplane0 = pass_major ? (grp * SH + sh) * S : si * S;
#pragma unroll
                    for (int sj = 0; sj < S; sj++) {
                        float c[4 + 4 * NSH];
                        c[0] = ACC(sh, sj, 0); c[1] = ACC(sh, sj, 1); c[2] = ACC(sh, sj, 2); c[3] = ACC(sh, sj, 3);
#pragma unroll
                        for (int d = 1; d <= NSH; d++)
#pragma unroll
                            for (int x = 0; x < 4; x++) c[4 * d + x] = __shfl_down_sync(0xffffffffu, c[x], d, 8);
                        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int v = 0; v < K; v++)
#pragma unroll
                            for (int x = 0; x < 4; x++) o[x] = fmaf(gk[v], c[x + v], o[x]);
                        *reinterpret_cast<float4 *>(cg + sj * plane) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(TG) : "memory");
#if UMPA_COLPASS4
                    // Column pass, FOUR output rows per thread.  The four threads of a strip column (chunk rows e4 ..
                    // e4+3: the four quarter warps of one warp) deal out the S planes -- thread qi takes planes qi, qi+4,
                    // ... -- and each filters its planes for all four rows: the 2Nw+4 row-filtered lines an output group
                    // needs are read ONCE per strip (one LDS.128 each) instead of once per output row: 2 instead of 5
                    // wavefronts per output line for Nw = 2.  FFMA2: the window factor is the broadcast scalar, the halves
                    // of a float4 line are natural register pairs.  E4 < 0: every tap lies inside this chunk (the
                    // common case, all addresses are immediates); E4 = 0, 4, 8: the group starts E4 rows into the
                    // chunk and its first taps reach into the previous chunk's last rows (carry) -- or, in a segment's
                    // first chunk, above the segment, where they feed no stored row.
                    if (any_store) {
                        float *dq = p.table + (size_t)(row0 + r_out4) * p.row_stride + (size_t)(si * S + qi) * p.plane_stride + tx0 + ec;
                        const float *cb = cbuf + (size_t)(grp * S + qi) * plane + ec;             // chunk row 0 of plane qi
                        const float *cr = carry + (size_t)((cplane0 + qi) * H2) * EXT_W + ec;     // carry row 0 of plane qi
                        auto colpass = [&](auto e4c) {
                            constexpr int E4 = decltype(e4c)::value;
                            const float *base = E4 < 0 ? cb + (e4 - H2) * EXT_W : cb;
#pragma unroll
                            for (int n = 0; n < (S + 3) / 4; n++) {
                                if (4 * n + 3 >= S && 4 * n + qi >= S) continue;           // (only the last n can run out of planes)
                                const float *bn = base + (4 * n) * plane;
                                float2 o[4][2];
#pragma unroll
                                for (int i = 0; i < 4; i++) o[i][0] = o[i][1] = make_float2(0.f, 0.f);
#pragma unroll
                                for (int t = 0; t < H2 + 4; t++) {                         // chunk row e4 - H2 + t
                                    float4 v;
                                    if (E4 < 0) v = *reinterpret_cast<const float4 *>(bn + t * EXT_W);
                                    else if (E4 + t >= H2) v = *reinterpret_cast<const float4 *>(bn + (E4 + t - H2) * EXT_W);
                                    else if (chunk > 0) v = *reinterpret_cast<const float4 *>(cr + ((4 * n) * H2 + E4 + t) * EXT_W);
                                    else v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                                    for (int i = 0; i < 4; i++) {
                                        const int u = t - i;                               // tap of output row i
                                        if (u >= 0 && u < K) {
                                            o[i][0] = __ffma2_rn(make_float2(gk[u], gk[u]), make_float2(v.x, v.y), o[i][0]);
                                            o[i][1] = __ffma2_rn(make_float2(gk[u], gk[u]), make_float2(v.z, v.w), o[i][1]);
                                        }
                                    }
                                }
#pragma unroll
                                for (int i = 0; i < 4; i++)
                                    if ((row_ok >> i) & 1)
                                        *reinterpret_cast<float4 *>(dq + (size_t)i * p.row_stride + (4 * n) * p.plane_stride) =
                                            make_float4(o[i][0].x, o[i][0].y, o[i][1].x, o[i][1].y);
                            }
                        };
                        if (e4 >= H2) colpass(std::integral_constant<int, -1>{});
                        else if (e4 == 0) colpass(std::integral_constant<int, 0>{});
                        else if (e4 == 4) colpass(std::integral_constant<int, (H2 > 4 ? 4 : 0)>{});
                        else colpass(std::integral_constant<int, (H2 > 8 ? 8 : 0)>{});
                    }
#else
                    // Column pass: each thread filters its own strip over the chunk rows er-2Nw .. er (FFMA2: the window
                    // factor is the broadcast scalar, the halves of a float4 line are natural register pairs)
                    if (store) {
                        float *dst = p.table + (size_t)(row0 + r_out) * p.row_stride + (size_t)(si * S) * p.plane_stride + tx0 + ec;
                        if (er >= H2) {                          // every tap inside this chunk
#pragma unroll
                            for (int sj = 0; sj < S; sj++) {
                                float2 o01 = make_float2(0.f, 0.f), o23 = make_float2(0.f, 0.f);
#pragma unroll
                                for (int u = 0; u < K; u++) {
                                    const float4 t = *reinterpret_cast<const float4 *>(cg + sj * plane + (u - H2) * EXT_W);
                                    o01 = __ffma2_rn(make_float2(gk[u], gk[u]), make_float2(t.x, t.y), o01);
                                    o23 = __ffma2_rn(make_float2(gk[u], gk[u]), make_float2(t.z, t.w), o23);
                                }
                                *reinterpret_cast<float4 *>(dst + sj * p.plane_stride) = make_float4(o01.x, o01.y, o23.x, o23.y);
                            }
                        } else {                                 // taps er+u < 2Nw come from the previous chunk's last rows
                            const float *cw = cbuf + (size_t)grp * S * plane + ec;
                            const float *cr = carry + (size_t)cplane0 * (H2 * EXT_W) + ec;
#pragma unroll
                            for (int sj = 0; sj < S; sj++) {
                                float2 o01 = make_float2(0.f, 0.f), o23 = make_float2(0.f, 0.f);
#pragma unroll
                                for (int u = 0; u < K; u++) {
                                    const int q = er + u;        // row of [carry rows | chunk rows]
                                    const float *src = q >= H2 ? cw + sj * plane + (q - H2) * EXT_W : cr + (sj * H2 + q) * EXT_W;
                                    const float4 t = *reinterpret_cast<const float4 *>(src);
                                    o01 = __ffma2_rn(make_float2(gk[u], gk[u]), make_float2(t.x, t.y), o01);
                                    o23 = __ffma2_rn(make_float2(gk[u], gk[u]), make_float2(t.z, t.w), o23);
                                }
                                *reinterpret_cast<float4 *>(dst + sj * p.plane_stride) = make_float4(o01.x, o01.y, o23.x, o23.y);
                            }
                        }
                    }
#endif
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(TG) : "memory");
                    if (H2 > 0 && keep) {                        // (after the barrier: the readers of the old carry are done)
                        float *cr = carry + ((size_t)cplane0 * H2 + (er - (p.EH - H2))) * EXT_W + ec;
#pragma unroll
                        for (int sj = 0; sj < S; sj++)
                            *reinterpret_cast<float4 *>(cr + sj * (H2 * EXT_W)) = *reinterpret_cast<const float4 *>(cg + sj * plane);
                    }
                }
            }
        }
      }
    }
#undef ACC
}

// compile-time choice of the per-thread shift-row block: keeps SH*S*4 accumulators <= ~120
template <int S> struct RowBlock { static constexpr int SH = S <= 9 ? 3 : (S <= 17 ? 2 : 1); };

template <int S, int NWT>
int launch_shift_table(const CUtensorMap &mapA, const CUtensorMap &mapB, const TableParams &p, dim3 grid, int nt,
                       size_t smem, cudaStream_t st)
{
    auto kern = shift_table_kernel<S, RowBlock<S>::SH, NWT>;
    static size_t attr_set[64] = {0};                // per device (function attributes belong to the device's context)
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { umpa_set_error("cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e)); return UMPA_ERR_CUDA; }
        attr_set[dev & 63] = smem;
    }
    kern<<<grid, nt, smem, st>>>(mapA, mapB, p);
    UMPA_CUDA(cudaGetLastError());
    return UMPA_OK;
}

template <int NWT>
int dispatch_shift_table_s(int S, const CUtensorMap &a, const CUtensorMap &b, const TableParams &p, dim3 grid, int nt,
                           size_t smem, cudaStream_t st)
{
    switch (S) {
        case 3: return launch_shift_table<3, NWT>(a, b, p, grid, nt, smem, st);
        case 5: return launch_shift_table<5, NWT>(a, b, p, grid, nt, smem, st);
        case 7: return launch_shift_table<7, NWT>(a, b, p, grid, nt, smem, st);
        case 9: return launch_shift_table<9, NWT>(a, b, p, grid, nt, smem, st);
        case 11: return launch_shift_table<11, NWT>(a, b, p, grid, nt, smem, st);
        case 13: return launch_shift_table<13, NWT>(a, b, p, grid, nt, smem, st);
        case 15: return launch_shift_table<15, NWT>(a, b, p, grid, nt, smem, st);
        case 17: return launch_shift_table<17, NWT>(a, b, p, grid, nt, smem, st);
        case 19: return launch_shift_table<19, NWT>(a, b, p, grid, nt, smem, st);
    }
    umpa_set_error("table path: max_shift %d not instantiated", (S + 1) / 2);
    return UMPA_ERR_UNSUPPORTED;
}

}  // namespace shift_table

// one entry point per translation unit of shift_table_inst.cu
#define UMPA_ST_ARGS int S, const CUtensorMap &a, const CUtensorMap &b, const shift_table::TableParams &p, dim3 grid, int nt, \
                     size_t smem, cudaStream_t st
int shift_table_launch_plain(UMPA_ST_ARGS);
int shift_table_launch_nw0(UMPA_ST_ARGS);
int shift_table_launch_nw1(UMPA_ST_ARGS);
int shift_table_launch_nw2(UMPA_ST_ARGS);
int shift_table_launch_nw3(UMPA_ST_ARGS);
int shift_table_launch_nw4(UMPA_ST_ARGS);
int shift_table_launch_nw5(UMPA_ST_ARGS);
int shift_table_launch_nw6(UMPA_ST_ARGS);
