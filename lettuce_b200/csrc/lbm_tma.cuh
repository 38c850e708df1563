// lbm_tma.cuh -- the fused time step with the streaming neighbourhood staged through shared memory by the
// Tensor Memory Accelerator (sm_100a: cp.async.bulk.tensor + mbarrier).
//
// Why: the LDG/STG step kernel (lbm_step.cuh) spends more than half of its issue slots on things that are not
// arithmetic -- two integer instructions per 64-bit address, the address-table loads, two 32-bit loads for every
// population that moves along the contiguous axis.  For the bandwidth-bound operators (BGK, TRT) that is free; the
// entropic KBC operator (~480 fp32 operations per node pair) is bound by the issue rate, not by HBM.  Here a
// persistent CTA per SM lets the TMA unit do ALL global addressing:
//
//   producer warps one thread per population, spread over four warps: per tile and population ONE bulk tensor load of
//                  the tile's rows, its box shifted by -e_q in x and y (pull: the gather of
//                  lettuce/_simulation.py:245-256 done by the copy engine), plus, for populations that move along z,
//                  one box with the quads beyond the rows' ends; later ONE bulk tensor store of the rows.  Tiles whose
//                  source rows wrap around the lattice fall back to one load per row (wrap folded into the
//                  coordinates) and 16-byte bulk copies of the quads
//   consumer warps two neighbouring nodes per thread as a float2 on the packed fp32 pipe: Q shared-memory loads
//                  with immediate offsets (the shift by one element along z happens here: a box must start on a
//                  16-byte boundary of global memory, measured with scripts/probes/tma_probe.cu), collide in
//                  registers, Q STS.64 back into the rows
//
// over a ring of NS shared-memory stages (full[] barriers: TMA bytes landed; done[] barriers: consumers' results are
// in shared memory).  The halo quad's address wraps around the contiguous axis like torch.roll does.
//
// step_tma_kernel takes the steps that do not push (PRE_STREAMING, NO_STREAMING), step_tma_push_kernel at the end of
// this file the pushing step (POST_STREAMING); fp32 lattices whose contiguous extent is a multiple of 64, the whole
// lattice on one GPU or the interior planes of a multi-GPU slab (pulling steps).  Everything else runs the LDG kernel.
// Results are bit-identical to it (same collide code, data movement only).
#pragma once
#include <cuda.h>

#include "lbm_step.cuh"

namespace lbm {

constexpr int kTmaConsumers = 256;                 // consumer threads per CTA, two nodes each
constexpr int kTmaProducers = 4;                   // producer warps (one issuing thread each)
constexpr int kTmaThreads = kTmaConsumers + 32 * kTmaProducers;
constexpr int kTmaTileNodes = 2 * kTmaConsumers;
constexpr int kTmaMaxStages = 8;
constexpr int kTmaMaxRows = 8;                     // tile rows = kTmaTileNodes / tz, tz >= 64

// tensor maps of the two population buffers, fp32 [Q][n0][n1][n2], built on the host (lbm_api.cu) and passed as one
// kernel parameter: `row` boxes are one tile row (tz, 1, 1, 1), `box` all rows of a tile (tz, rows, 1, 1) -- (tz, 1,
// rows, 1) for 2-D lattices, whose rows are consecutive in x --, `halo` the quads beyond the rows' ends (4, rows, 1, 1)
struct TmaMaps {
    CUtensorMap in_row, in_box, in_halo, out_row, out_box;
};

struct TmaParams {
    const float *in;
    float *out;
    unsigned *counters;       // [0] next tile to hand out, [1] CTAs that have run out of tiles; both 0 between launches
    int64_t N;                // n0 * n1 * n2
    int n0, n1, n2;
    int tz, tz_log2;          // z extent of a tile row (box width), a power of two dividing n2
    int rows;                 // tile rows: kTmaTileNodes / tz consecutive (x, y) rows
    int zchunks;              // n2 / tz
    int row_begin, n_rows;    // rows (x * n1 + y) this launch covers: [row_begin, n_rows); the whole lattice, or
                              // the interior planes of a multi-GPU slab (its cut planes need the neighbours)
    int n_tiles;
    int skip_wait;            // launched behind a kernel of the SAME step (slab: the cut-plane kernel): do not wait
                              // for that grid to complete, its start already implies that the previous step is done
    double *partials;         // REDUCE instantiation: one (sum 0.5|u|^2, max |u|^2) pair per CTA of the state the step
    int reduce_slots;         // writes -- sums in [0, reduce_slots), maxima in [reduce_slots, 2 reduce_slots)
    int stages;
    int boxable;              // a full tile never straddles two planes (3-D) / rows divide n0 (2-D): box maps usable
    int reverse;              // sweep the tiles backwards (L2 reuse between consecutive steps)
    float ca, cb;
    ForceArgs<float> force;
};

// e_q as run-time data for the producer lanes (S::e is a compile-time table; indexed with a run-time q it would be
// rebuilt on every thread's stack)
static __constant__ signed char kVelocityTable[3][27][3] = {
    {{0, 0, 0}, {1, 0, 0}, {0, 0, 1}, {-1, 0, 0}, {0, 0, -1}, {1, 0, 1}, {-1, 0, 1}, {-1, 0, -1}, {1, 0, -1}},
    {{0, 0, 0},  {1, 0, 0},   {-1, 0, 0}, {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1},  {0, 1, 1},  {0, -1, -1},
     {0, 1, -1}, {0, -1, 1},  {1, 0, 1},  {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1}, {1, 1, 0},  {-1, -1, 0}, {1, -1, 0},
     {-1, 1, 0}},
    {{0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},   {0, -1, 0},  {0, 0, 1},   {0, 0, -1},  {0, 1, 1},   {0, -1, -1},
     {0, 1, -1},  {0, -1, 1},  {1, 0, 1},   {-1, 0, -1}, {1, 0, -1},  {-1, 0, 1},  {1, 1, 0},   {-1, -1, 0}, {1, -1, 0},
     {-1, 1, 0},  {1, 1, 1},   {-1, -1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1},  {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1}}};
template <class S>
LBM_D int velocity_component(int q, int a) {
    constexpr int row = S::Q == 9 ? 0 : (S::Q == 19 ? 1 : 2);
    return kVelocityTable[row][q][a];
}
// the table above restates S::e
template <class S>
constexpr bool velocity_table_matches(const signed char (&t)[27][3]) {
    for (int q = 0; q < S::Q; ++q)
        for (int a = 0; a < 3; ++a)
            if (t[q][a] != S::e(q, a)) return false;
    return true;
}

// A stage: Q populations of kTmaTileNodes floats, then one 128-byte slot (a quad per tile row) for every population
// that moves along z.  (TMA needs 128-byte aligned shared-memory boxes: scripts/probes/tma_probe.cu, modes 10 / 11.)
constexpr int kTmaHaloSlot = kTmaMaxRows * 4;
template <class S>
constexpr int tma_z_populations() {
    int n = 0;
    for (int q = 0; q < S::Q; ++q) n += S::e(q, 2) != 0;
    return n;
}
// index of population q among those that move along z
template <class S>
constexpr int tma_z_slot(int q) {
    int n = 0;
    for (int k = 0; k < q; ++k) n += S::e(k, 2) != 0;
    return n;
}
template <class S>
constexpr int tma_stage_floats() { return S::Q * kTmaTileNodes + tma_z_populations<S>() * kTmaHaloSlot; }
template <class S>
constexpr size_t tma_smem_bytes(int stages) { return (size_t)stages * tma_stage_floats<S>() * sizeof(float); }

namespace tma {

LBM_D uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

LBM_D void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
LBM_D void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
LBM_D void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
LBM_D void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// global -> shared, 16 bytes (both addresses 16-byte aligned), completion counted on `bar`
LBM_D void load_16(uint32_t dst, const void *src, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                 ::"r"(dst), "l"(src), "r"(bar)
                 : "memory");
}
// global -> shared: box of `map` at (c0, c1, c2, c3), completion counted in bytes on `bar`
LBM_D void load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
// shared -> global: box of `map` at (c0, c1, c2, c3); elements outside the tensor are not written
LBM_D void store_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
LBM_D void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest bulk group of this thread have finished READING shared memory
LBM_D void store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
LBM_D void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
LBM_D void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tma

// Tiles are handed out dynamically: SMs do not get the same share of the memory system (with a static round-robin
// split ncu showed 912 k to 1106 k active cycles per SM, the slowest one setting the kernel's duration).  One producer
// thread per CTA claims tile numbers from a global counter -- a few tiles ahead of their use, so that the atomic's
// latency stays hidden --, works out the tile's geometry once and publishes it to the CTA through a small ring in
// shared memory.  The last CTA to run out of tiles resets the counters for the next launch.
constexpr int kTmaRing = 2 * kTmaMaxStages;     // claimed-tile ring: never more than stages + 1 entries in use
struct TileInfo {
    int z0, r0, x0, y0;                         // first z, first row index (-1: no more tiles), its (x, y)
};

// CTAs per SM the kernel is compiled for: two for the small velocity sets (their tiles are short, a second CTA keeps
// the SM busy while the first one waits), one for D3Q27 (168 registers, three 56 KB stages)
template <class S>
constexpr int tma_ctas_per_sm() { return S::Q <= 19 ? 2 : 1; }

template <class S, int COLL, bool PULL, bool REDUCE>
__global__ void __launch_bounds__(kTmaThreads, tma_ctas_per_sm<S>())
    step_tma_kernel(const __grid_constant__ TmaMaps maps, const __grid_constant__ TmaParams p) {
    constexpr int Q = S::Q;
    constexpr int T = kTmaTileNodes;
    constexpr int kStageFloats = tma_stage_floats<S>();
    // (TMA wants 128-byte aligned shared-memory boxes; the declared alignment places the dynamic window)
    extern __shared__ __align__(128) float stage0[];
    __shared__ __align__(8) unsigned long long full_bar[kTmaMaxStages], done_bar[kTmaMaxStages], ring_bar[kTmaRing];
    __shared__ TileInfo ring[kTmaRing];
    __shared__ double reduce_part[2][kTmaConsumers / 32];
    if (threadIdx.x == 0 && (tma::smem_addr(stage0) & 127u)) __trap();

    // (see step_kernel: complete the step in front, then let the kernel behind become resident)
    if (!p.skip_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int NS = p.stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            tma::mbar_init(tma::smem_addr(&full_bar[s]), Q);               // one arrive.expect_tx per producer thread
            tma::mbar_init(tma::smem_addr(&done_bar[s]), kTmaConsumers);   // every consumer thread arrives
        }
        for (int s = 0; s < kTmaRing; ++s) tma::mbar_init(tma::smem_addr(&ring_bar[s]), 1);   // the claiming thread
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma::fence_async_shared();
    }
    __syncthreads();

    // the i-th tile of this CTA, once the claiming thread has published it
    auto tile = [&](int i) {
        tma::mbar_wait(tma::smem_addr(&ring_bar[i % kTmaRing]), (uint32_t)(i / kTmaRing) & 1u);
        return ring[i % kTmaRing];
    };

    if (threadIdx.x >= kTmaConsumers) {
        // ---------------------------------- producer warps: one thread per population (lane l of warp w: q = w + P l)
        // The copy instructions take their operands from uniform registers, so the threads of ONE warp issue one after
        // the other; the address arithmetic in front of them runs in parallel, and the P warps issue concurrently.
        const int q = ((threadIdx.x - kTmaConsumers) >> 5) + kTmaProducers * (threadIdx.x & 31);
        if (q >= Q) return;
        const int e0 = PULL ? velocity_component<S>(q, 0) : 0, e1 = PULL ? velocity_component<S>(q, 1) : 0,
                  e2 = PULL ? velocity_component<S>(q, 2) : 0;
        int zslot = 0;                                    // index of q among the populations that move along z
        for (int k = 0; k < q; ++k) zslot += velocity_component<S>(k, 2) != 0;
        const uint32_t row_bytes = (uint32_t)p.tz * sizeof(float);
        // rows of a tile run along y (3-D lattices) or along x (2-D lattices, n1 = 1)
        const bool rows_along_y = p.n1 > 1;

        // the thread of population 0 claims the tiles; `claimed` counts the ring entries it has published
        const bool claimer = q == 0;
        int claimed = 0;
        bool exhausted = false;
        auto publish = [&](unsigned k) {
            TileInfo info;
            if (k >= (unsigned)p.n_tiles) {
                info.z0 = info.x0 = info.y0 = 0;
                info.r0 = -1;
                if (!exhausted) {
                    exhausted = true;
                    // this CTA will not touch the counter again; the last CTA to get here rearms it
                    if (atomicAdd(p.counters + 1, 1u) == gridDim.x - 1) {
                        p.counters[0] = 0;
                        p.counters[1] = 0;
                    }
                }
            } else {
                const int t = p.reverse ? p.n_tiles - 1 - (int)k : (int)k;
                const int zc = t % p.zchunks;
                info.z0 = zc << p.tz_log2;
                info.r0 = p.row_begin + (t / p.zchunks) * p.rows;
                info.x0 = info.r0 / p.n1;
                info.y0 = info.r0 - info.x0 * p.n1;
            }
            ring[claimed % kTmaRing] = info;
            tma::mbar_arrive(tma::smem_addr(&ring_bar[claimed % kTmaRing]));     // (release: the entry is visible)
            ++claimed;
        };
        if (claimer) {
            // the first NS tiles with one atomic: NS - 1 for the prologue, one ahead
            const unsigned k0 = atomicAdd(p.counters, (unsigned)NS);
            for (int i = 0; i < NS; ++i) publish(exhausted ? 0xffffffffu : k0 + i);
        }

        auto issue_loads = [&](int i, const TileInfo &ti) {
            const int s = i % NS;
            const int z0 = ti.z0, r0 = ti.r0, x0 = ti.x0, y0 = ti.y0;
            const int rows = min(p.rows, p.n_rows - r0);
            const uint32_t bar = tma::smem_addr(&full_bar[s]);
            float *pop = stage0 + (size_t)s * kStageFloats + q * T;
            float *halo = stage0 + (size_t)s * kStageFloats + Q * T + zslot * kTmaHaloSlot;
            tma::mbar_arrive_expect_tx(bar, (uint32_t)rows * (row_bytes + (e2 != 0 ? 16u : 0u)));
            // the quad that holds the neighbour across the row's end along z (periodic): z0 - 4 .. z0 - 1 for
            // populations that move towards +z, z0 + tz .. z0 + tz + 3 for those that move towards -z
            int zq = e2 == 1 ? z0 - 4 : z0 + p.tz;
            zq = zq < 0 ? zq + p.n2 : (zq >= p.n2 ? zq - p.n2 : zq);
            // all rows of the tile as ONE box when they are consecutive in the source as well (no wrap between them)
            const int xs0 = x0 - e0, ys0 = y0 - e1;
            const bool box = p.boxable && rows == p.rows &&
                             (rows_along_y ? (ys0 >= 0 && ys0 + rows <= p.n1) : (xs0 >= 0 && xs0 + rows <= p.n0));
            if (box) {
                const int xs = rows_along_y ? wrap(xs0, p.n0) : xs0;
                // (the rows themselves at their own z: a box has to start on a 16-byte boundary of global memory,
                // the shift by one element along z is done by the consumers' shared-memory reads)
                tma::load_4d(tma::smem_addr(pop), &maps.in_box, z0, ys0, xs, q, bar);
                if (e2 != 0) tma::load_4d(tma::smem_addr(halo), &maps.in_halo, zq, ys0, xs, q, bar);
            } else {
                for (int j = 0; j < rows; ++j) {
                    const int r = r0 + j;
                    const int x = r / p.n1, y = r - x * p.n1;
                    const int xs = wrap(x - e0, p.n0), ys = wrap(y - e1, p.n1);
                    tma::load_4d(tma::smem_addr(pop + (j << p.tz_log2)), &maps.in_row, z0, ys, xs, q, bar);
                    if (e2 != 0)
                        tma::load_16(tma::smem_addr(halo + j * 4),
                                     p.in + q * p.N + ((int64_t)xs * p.n1 + ys) * p.n2 + zq, bar);
                }
            }
        };
        auto issue_stores = [&](int i, const TileInfo &ti) {
            const int s = i % NS;
            const int rows = min(p.rows, p.n_rows - ti.r0);
            float *pop = stage0 + (size_t)s * kStageFloats + q * T;
            if (p.boxable && rows == p.rows) {
                tma::store_4d(&maps.out_box, ti.z0, ti.y0, ti.x0, q, tma::smem_addr(pop));
            } else {
                for (int j = 0; j < rows; ++j) {
                    const int r = ti.r0 + j;
                    const int x = r / p.n1, y = r - x * p.n1;
                    tma::store_4d(&maps.out_row, ti.z0, y, x, q, tma::smem_addr(pop + (j << p.tz_log2)));
                }
            }
            tma::store_commit();
        };
        int loaded = 0;                 // tiles whose loads have been issued
        bool more = true;               // the ring has not shown its end marker yet
        for (int i = 0; i < NS - 1 && more; ++i) {
            const TileInfo ti = tile(i);
            if (ti.r0 < 0) more = false;
            else { issue_loads(i, ti); ++loaded; }
        }
        for (int i = 0; i < loaded; ++i) {
            const int s = i % NS;
            const TileInfo ti = ring[i % kTmaRing];            // (read before: still in the ring, kTmaRing >= 2 NS)
            tma::mbar_wait(tma::smem_addr(&done_bar[s]), (uint32_t)(i / NS) & 1u);
            issue_stores(i, ti);
            if (more) {
                const int n = i + NS - 1;
                const TileInfo tn = tile(n);
                if (tn.r0 < 0) {
                    more = false;
                } else {
                    // the stage of tile i-1 is the next to be filled: its store (this thread's previous bulk group)
                    // has to be done reading the rows of population q first
                    tma::store_wait_read_1();
                    issue_loads(n, tn);
                    ++loaded;
                    // one more tile for the iteration after this one
                    if (claimer) publish(exhausted ? 0xffffffffu : atomicAdd(p.counters, 1u));
                }
            }
        }
        tma::store_wait_all();
        return;
    }

    // ---------------------------------------------------------------------- consumers: two nodes per thread
    const int t = threadIdx.x;
    const int j = (2 * t) >> p.tz_log2;              // tile row
    const int zl = (2 * t) & (p.tz - 1);             // first of the two nodes within the row
    // The pair itself sits at float 2 t of its population; the elements in front of it (z - 1) and behind it (z + 2)
    // are its neighbours in the row, except at the row's ends, where they come from the halo quad.
    const int i0 = 2 * t;
    const bool first = zl == 0, last = zl == p.tz - 2;
    double e_sum = 0.0, e_max = 0.0;                 // REDUCE: moments of the nodes this thread has written
    for (int i = 0;; ++i) {
        const int s = i % NS;
        const TileInfo ti = tile(i);
        if (ti.r0 < 0) break;
        float *st = stage0 + (size_t)s * kStageFloats;
        tma::mbar_wait(tma::smem_addr(&full_bar[s]), (uint32_t)(i / NS) & 1u);
        float2 f[Q];
        const bool valid = ti.r0 + j < p.n_rows;
        if (valid) {
            ForQ<Q>::run([&]<int q>() {
                constexpr int e2 = PULL ? S::e(q, 2) : 0;
                const float *pop = st + q * T + i0;
                if constexpr (e2 == 0) f[q] = *reinterpret_cast<const float2 *>(pop);
                else if constexpr (e2 == 1) f[q] = make_float2(pop[-1], pop[0]);             // from z - 1, z
                else f[q] = make_float2(pop[1], pop[2]);                                      // from z + 1, z + 2
            });
            if constexpr (PULL) {
                const float *halo = st + Q * T + j * 4;
                if (first) {
                    ForQ<Q>::run([&]<int q>() {
                        if constexpr (S::e(q, 2) == 1) f[q].x = halo[tma_z_slot<S>(q) * kTmaHaloSlot + 3];
                    });
                }
                if (last) {
                    ForQ<Q>::run([&]<int q>() {
                        if constexpr (S::e(q, 2) == -1) f[q].y = halo[tma_z_slot<S>(q) * kTmaHaloSlot];
                    });
                }
            }
        }
        // The results go back into the slots the inputs came from, and a pair's neighbours along z belong to other
        // threads: nobody may store before everybody has loaded (without this: single wrong nodes at warp boundaries).
        if constexpr (PULL) asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumers) : "memory");
        if (valid) {
            collide_lanes<S, float2, COLL>(p, f);
            float2 *out = reinterpret_cast<float2 *>(st + i0);
            ForQ<Q>::run([&]<int q>() { out[q * (T / 2)] = f[q]; });
            if constexpr (REDUCE) {
                const bool take[2] = {true, true};
                accumulate_kinetic<S, float2>(f, take, e_sum, e_max);
            }
        }
        tma::fence_async_shared();
        tma::mbar_arrive(tma::smem_addr(&done_bar[s]));
    }
    if constexpr (REDUCE) {
        // one (sum, max) pair per CTA, folded in a fixed order among the consumer warps (the producers are elsewhere:
        // a named barrier for the 256 consumers)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            e_sum += __shfl_down_sync(0xffffffffu, e_sum, o);
            e_max = fmax(e_max, __shfl_down_sync(0xffffffffu, e_max, o));
        }
        if ((t & 31) == 0) {
            reduce_part[0][t >> 5] = e_sum;
            reduce_part[1][t >> 5] = e_max;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(kTmaConsumers) : "memory");
        if (t == 0) {
            double sum = 0.0, mx = 0.0;
            for (int w = 0; w < kTmaConsumers / 32; ++w) {
                sum += reduce_part[0][w];
                mx = fmax(mx, reduce_part[1][w]);
            }
            p.partials[blockIdx.x] = sum;
            p.partials[p.reduce_slots + blockIdx.x] = mx;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The pushing step (POST_STREAMING, the default of lettuce's Simulation) on the same machinery.
//
// A bulk tensor store cannot write a row shifted by one element along z either (16-byte box starts), and the element
// that would arrive at the row's first / last position comes from a node of the NEIGHBOURING tile.  So every tile
// also loads the quads next to its rows' ends -- for ALL populations -- and a ninth consumer warp ("halo warp")
// collides those 2 x rows neighbour nodes itself, one node per lane, next to the eight warps that do the tile's own
// nodes two per thread: overlapped tiling, ~1 % redundant arithmetic and bytes.  The consumers then write their results
// into the rows shifted by e_z in shared memory (position k of a row = what arrives at z0 + k), and the producers store
// every row with one aligned box at its destination (x + e_x, y + e_y, z0): no scalar stores, no second pass.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTmaPushProducers = 3;                       // producer warps of the pushing kernel (384 threads in all)
constexpr int kTmaPushConsumers = kTmaConsumers + 32;      // + the halo warp
template <class S>
constexpr int tma_push_stage_floats() { return S::Q * (kTmaTileNodes + 2 * kTmaHaloSlot); }
template <class S>
constexpr size_t tma_push_smem_bytes(int stages) { return (size_t)stages * tma_push_stage_floats<S>() * sizeof(float); }

template <class S, int COLL>
__global__ void __launch_bounds__(kTmaThreads, tma_ctas_per_sm<S>())
    step_tma_push_kernel(const __grid_constant__ TmaMaps maps, const __grid_constant__ TmaParams p) {
    constexpr int Q = S::Q;
    constexpr int T = kTmaTileNodes;
    constexpr int kStageFloats = tma_push_stage_floats<S>();
    static_assert(kTmaPushConsumers + 32 * kTmaPushProducers == kTmaThreads, "thread roles");
    extern __shared__ __align__(128) float stage0[];
    __shared__ __align__(8) unsigned long long full_bar[kTmaMaxStages], done_bar[kTmaMaxStages], ring_bar[kTmaRing];
    __shared__ TileInfo ring[kTmaRing];
    if (threadIdx.x == 0 && (tma::smem_addr(stage0) & 127u)) __trap();

    if (!p.skip_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int NS = p.stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            tma::mbar_init(tma::smem_addr(&full_bar[s]), Q);
            tma::mbar_init(tma::smem_addr(&done_bar[s]), kTmaPushConsumers);
        }
        for (int s = 0; s < kTmaRing; ++s) tma::mbar_init(tma::smem_addr(&ring_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma::fence_async_shared();
    }
    __syncthreads();

    auto tile = [&](int i) {
        tma::mbar_wait(tma::smem_addr(&ring_bar[i % kTmaRing]), (uint32_t)(i / kTmaRing) & 1u);
        return ring[i % kTmaRing];
    };
    // halo quads of a stage: per population a slot for the quads in front of the rows (lo: z0 - 4 .. z0 - 1) and one
    // for the quads behind them (hi: z0 + tz .. z0 + tz + 3), a quad per tile row each
    auto halo_of = [&](float *st, int q, int side) { return st + Q * T + (q * 2 + side) * kTmaHaloSlot; };

    if (threadIdx.x >= kTmaPushConsumers) {
        // ------------------------------------------------------------------ producers: one thread per population
        const int pt = threadIdx.x - kTmaPushConsumers;
        const int q = (pt >> 5) + kTmaPushProducers * (pt & 31);
        if (q >= Q) return;
        const int e0 = velocity_component<S>(q, 0), e1 = velocity_component<S>(q, 1);
        const uint32_t row_bytes = (uint32_t)p.tz * sizeof(float);
        const bool rows_along_y = p.n1 > 1;
        const bool claimer = q == 0;
        int claimed = 0;
        bool exhausted = false;
        auto publish = [&](unsigned k) {
            TileInfo info;
            if (k >= (unsigned)p.n_tiles) {
                info.z0 = info.x0 = info.y0 = 0;
                info.r0 = -1;
                if (!exhausted) {
                    exhausted = true;
                    if (atomicAdd(p.counters + 1, 1u) == gridDim.x - 1) {
                        p.counters[0] = 0;
                        p.counters[1] = 0;
                    }
                }
            } else {
                const int t = p.reverse ? p.n_tiles - 1 - (int)k : (int)k;
                const int zc = t % p.zchunks;
                info.z0 = zc << p.tz_log2;
                info.r0 = p.row_begin + (t / p.zchunks) * p.rows;
                info.x0 = info.r0 / p.n1;
                info.y0 = info.r0 - info.x0 * p.n1;
            }
            ring[claimed % kTmaRing] = info;
            tma::mbar_arrive(tma::smem_addr(&ring_bar[claimed % kTmaRing]));
            ++claimed;
        };
        if (claimer) {
            const unsigned k0 = atomicAdd(p.counters, (unsigned)NS);
            for (int i = 0; i < NS; ++i) publish(exhausted ? 0xffffffffu : k0 + i);
        }
        auto issue_loads = [&](int i, const TileInfo &ti) {
            const int s = i % NS;
            const int rows = min(p.rows, p.n_rows - ti.r0);
            const uint32_t bar = tma::smem_addr(&full_bar[s]);
            float *st = stage0 + (size_t)s * kStageFloats;
            float *pop = st + q * T;
            tma::mbar_arrive_expect_tx(bar, (uint32_t)rows * (row_bytes + 32u));
            int zlo = ti.z0 - 4, zhi = ti.z0 + p.tz;             // the neighbours across the rows' ends, periodic
            if (zlo < 0) zlo += p.n2;
            if (zhi >= p.n2) zhi -= p.n2;
            if (p.boxable && rows == p.rows) {                   // (the rows at their own place: never a wrap)
                tma::load_4d(tma::smem_addr(pop), &maps.in_box, ti.z0, ti.y0, ti.x0, q, bar);
                tma::load_4d(tma::smem_addr(halo_of(st, q, 0)), &maps.in_halo, zlo, ti.y0, ti.x0, q, bar);
                tma::load_4d(tma::smem_addr(halo_of(st, q, 1)), &maps.in_halo, zhi, ti.y0, ti.x0, q, bar);
            } else {
                for (int j = 0; j < rows; ++j) {
                    const int r = ti.r0 + j;
                    const int x = r / p.n1, y = r - x * p.n1;
                    const float *src = p.in + q * p.N + ((int64_t)x * p.n1 + y) * p.n2;
                    tma::load_4d(tma::smem_addr(pop + (j << p.tz_log2)), &maps.in_row, ti.z0, y, x, q, bar);
                    tma::load_16(tma::smem_addr(halo_of(st, q, 0) + j * 4), src + zlo, bar);
                    tma::load_16(tma::smem_addr(halo_of(st, q, 1) + j * 4), src + zhi, bar);
                }
            }
        };
        auto issue_stores = [&](int i, const TileInfo &ti) {
            const int s = i % NS;
            const int rows = min(p.rows, p.n_rows - ti.r0);
            float *pop = stage0 + (size_t)s * kStageFloats + q * T;
            // the rows arrive at (x + e_x, y + e_y), still consecutive unless they wrap around the lattice
            const int xd0 = ti.x0 + e0, yd0 = ti.y0 + e1;
            const bool box = p.boxable && rows == p.rows &&
                             (rows_along_y ? (yd0 >= 0 && yd0 + rows <= p.n1) : (xd0 >= 0 && xd0 + rows <= p.n0));
            if (box) {
                tma::store_4d(&maps.out_box, ti.z0, yd0, rows_along_y ? wrap(xd0, p.n0) : xd0, q, tma::smem_addr(pop));
            } else {
                for (int j = 0; j < rows; ++j) {
                    const int r = ti.r0 + j;
                    const int x = r / p.n1, y = r - x * p.n1;
                    tma::store_4d(&maps.out_row, ti.z0, wrap(y + e1, p.n1), wrap(x + e0, p.n0), q,
                                  tma::smem_addr(pop + (j << p.tz_log2)));
                }
            }
            tma::store_commit();
        };
        int loaded = 0;
        bool more = true;
        for (int i = 0; i < NS - 1 && more; ++i) {
            const TileInfo ti = tile(i);
            if (ti.r0 < 0) more = false;
            else { issue_loads(i, ti); ++loaded; }
        }
        for (int i = 0; i < loaded; ++i) {
            const int s = i % NS;
            const TileInfo ti = ring[i % kTmaRing];
            tma::mbar_wait(tma::smem_addr(&done_bar[s]), (uint32_t)(i / NS) & 1u);
            issue_stores(i, ti);
            if (more) {
                const int n = i + NS - 1;
                const TileInfo tn = tile(n);
                if (tn.r0 < 0) {
                    more = false;
                } else {
                    tma::store_wait_read_1();
                    issue_loads(n, tn);
                    ++loaded;
                    if (claimer) publish(exhausted ? 0xffffffffu : atomicAdd(p.counters, 1u));
                }
            }
        }
        tma::store_wait_all();
        return;
    }

    const int t = threadIdx.x;
    if (t < kTmaConsumers) {
        // -------------------------------------------------------------- the tile's own nodes, two per thread
        const int j = (2 * t) >> p.tz_log2;
        const int zl = (2 * t) & (p.tz - 1);
        const int i0 = 2 * t;
        const bool first = zl == 0, last = zl == p.tz - 2;
        for (int i = 0;; ++i) {
            const int s = i % NS;
            const TileInfo ti = tile(i);
            if (ti.r0 < 0) break;
            float *st = stage0 + (size_t)s * kStageFloats;
            tma::mbar_wait(tma::smem_addr(&full_bar[s]), (uint32_t)(i / NS) & 1u);
            float2 f[Q];
            const bool valid = ti.r0 + j < p.n_rows;
            if (valid) {
                ForQ<Q>::run([&]<int q>() { f[q] = *reinterpret_cast<const float2 *>(st + q * T + i0); });
            }
            // (results are written into the rows the inputs came from, shifted: nobody stores before everybody has loaded)
            asm volatile("bar.sync 1, %0;" ::"n"(kTmaPushConsumers) : "memory");
            if (valid) {
                collide_lanes<S, float2, COLL>(p, f);
                ForQ<Q>::run([&]<int q>() {
                    constexpr int e2 = S::e(q, 2);
                    float *pop = st + q * T + i0;
                    if constexpr (e2 == 0) {
                        *reinterpret_cast<float2 *>(pop) = f[q];
                    } else if constexpr (e2 == 1) {            // arrives at z + 1, z + 2
                        pop[1] = f[q].x;
                        if (!last) pop[2] = f[q].y;            // (the row's last node feeds the next tile: its halo warp)
                    } else {                                   // arrives at z - 1, z
                        if (!first) pop[-1] = f[q].x;
                        pop[0] = f[q].y;
                    }
                });
            }
            tma::fence_async_shared();
            tma::mbar_arrive(tma::smem_addr(&done_bar[s]));
        }
    } else {
        // -------------------------------------------------------------- halo warp: the nodes beyond the rows' ends
        const int lane = t - kTmaConsumers;
        const int j = lane >> 1, side = lane & 1;      // row; 0: node z0 - 1 (feeds position 0), 1: node z0 + tz
        for (int i = 0;; ++i) {
            const int s = i % NS;
            const TileInfo ti = tile(i);
            if (ti.r0 < 0) break;
            float *st = stage0 + (size_t)s * kStageFloats;
            tma::mbar_wait(tma::smem_addr(&full_bar[s]), (uint32_t)(i / NS) & 1u);
            float g[Q];
            const bool valid = j < p.rows && ti.r0 + j < p.n_rows;
            if (valid) {
                ForQ<Q>::run([&]<int q>() { g[q] = halo_of(st, q, side)[j * 4 + (side ? 0 : 3)]; });
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTmaPushConsumers) : "memory");
            if (valid) {
                collide_lanes<S, float, COLL>(p, g);
                float *row = st + (j << p.tz_log2);
                ForQ<Q>::run([&]<int q>() {
                    constexpr int e2 = S::e(q, 2);
                    if constexpr (e2 == 1) {
                        if (side == 0) row[q * T] = g[q];                    // node z0 - 1 arrives at z0
                    } else if constexpr (e2 == -1) {
                        if (side == 1) row[q * T + p.tz - 1] = g[q];         // node z0 + tz arrives at z0 + tz - 1
                    }
                });
            }
            tma::fence_async_shared();
            tma::mbar_arrive(tma::smem_addr(&done_bar[s]));
        }
    }
}

}  // namespace lbm
