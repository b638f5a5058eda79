// vdbrt_kernels.cuh -- the sm_100a kernels of the ray-tracing hot path (included once, by vdbrt.cu).
//
// Scheduling (subsystem (a) of the north star): persistent CTAs, one per SM x resident CTAs, whose warps pull
// 8x4-pixel tiles from a global atomic work queue.  Tiles are numbered macro-tile major (default 64x64 pixels)
// so that consecutive queue tickets touch neighbouring rays -> neighbouring tree nodes (L1/L2 reuse), and so that
// a multi-GPU partition (vdbrt_partition) is a stride over macro tiles.
#pragma once
#include "vdbrt_device.cuh"

namespace vdbrt {

constexpr int kBlockThreads = 128;       // 4 warps per CTA
#ifndef VDBRT_SUBW
#define VDBRT_SUBW 8
#endif
constexpr int kSubW = VDBRT_SUBW, kSubH = 32 / VDBRT_SUBW;      // pixels per warp tile (8 x 4; 4 x 8 and 16 x 2 were measured, see profiles/r02_summary.md)

struct TileMap {
    uint32_t width, height;
    uint32_t tile_w, tile_h;             // macro tile
    uint32_t macro_x, macro_count;       // macro tiles per row / total
    uint32_t rank, count;                // partition
    uint32_t sub_x, sub_per_macro;       // warp tiles per macro-tile row / per macro tile
    uint32_t items;                      // work items (warp tiles) owned by this rank
};

__host__ inline TileMap makeTileMap(uint32_t width, uint32_t height, uint32_t tw, uint32_t th, uint32_t rank, uint32_t count)
{
    TileMap m;
    m.width = width; m.height = height;
    m.tile_w = tw ? tw : 64u; m.tile_h = th ? th : 64u;
    m.count = count ? count : 1u; m.rank = count ? rank : 0u;
    m.macro_x = (width + m.tile_w - 1) / m.tile_w;
    const uint32_t macro_y = (height + m.tile_h - 1) / m.tile_h;
    m.macro_count = m.macro_x * macro_y;
    m.sub_x = (m.tile_w + kSubW - 1) / kSubW;
    m.sub_per_macro = m.sub_x * ((m.tile_h + kSubH - 1) / kSubH);
    const uint32_t owned = m.macro_count > m.rank ? (m.macro_count - m.rank + m.count - 1) / m.count : 0u;
    m.items = owned * m.sub_per_macro;
    return m;
}

// warp-level ticket: lane 0 takes the next work item, everyone gets the pixel of its lane (or false)
__device__ __forceinline__ bool nextPixel(const TileMap& m, unsigned int* queue, uint32_t& px, uint32_t& py, bool& valid)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(queue, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= m.items) return false;
    const uint32_t macro = m.rank + (item / m.sub_per_macro) * m.count;
    const uint32_t sub = item % m.sub_per_macro;
    const uint32_t mx = macro % m.macro_x, my = macro / m.macro_x;
    const uint32_t lx = (sub % m.sub_x) * kSubW + (lane % kSubW), ly = (sub / m.sub_x) * kSubH + (lane / kSubW);
    px = mx * m.tile_w + lx; py = my * m.tile_h + ly;
    valid = lx < m.tile_w && ly < m.tile_h && px < m.width && py < m.height;
    return true;
}

struct AuxOut { uint8_t* hit; int32_t* ijk; double* t_index; double* t_world; double* xyz; double* nml; };
struct LsParams { float iso, vmin, vmax, frac; uint32_t sub; uint32_t uniform_bg; float bg[4]; uint32_t iters, pad; double jitter[16];
                  const float4* bg_film; };   // where a miss reads its old pixel from (the film itself, or the device copy of a host film)

__device__ __forceinline__ void flushCounters(const Counters& c, unsigned long long* out)
{
    // warp-reduce then one atomic per counter per warp
    const uint32_t vals[10] = {c.rays, c.root, c.upper, c.lower, c.voxel, c.refills, c.psamples, c.ssamples, c.srays, c.hits};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        uint32_t v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(out + k, (unsigned long long)v);
    }
}

__device__ __forceinline__ void flushDiag(const Counters& c, unsigned long long* out)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t v = c.diag[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(out + k, (unsigned long long)v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// LevelSetRayTracer::operator() (tools/RayTracer.h:899-918): primary ray + (spp-1) jittered rays per pixel.
// Jitter index n(i,j) = 2*(spp-1)*(j*W+i): the reference's counter for render(threaded=false).
//
// Warp-synchronous persistent loop: every lane owns one pixel at a time and advances its ray by ONE traversal step
// per iteration (lsAdvance); the warp reconverges at the top of every iteration.  Lanes whose pixel is finished
// are re-fed from the atomic queue in batches (one atomicAdd per refill, tickets numbered in tile order so a warp's
// lanes stay spatially coherent), instead of idling until the slowest ray of a fixed tile is done.
// ------------------------------------------------------------------------------------------------------------
// CTAs per SM.  5 (96 registers; what ptxas then spills is the pixel bookkeeping and the hit record, touched once per ray) against 4
// (122 registers, no spills): C4 34.1 -> 32.0 ms, C2 2.68 -> 2.60 ms; 6 (80 registers) spills inside the traversal and loses (36.5 / 3.1 ms).
// The LONG instantiations (small shares of a frame, bounded by their slowest tile rather than by throughput) keep 4: C2 1/8 0.60 against 0.76 ms.
#ifndef VDBRT_MINBLOCKS
#define VDBRT_MINBLOCKS 5
#endif
#ifndef VDBRT_SHADE_AT_END
#define VDBRT_SHADE_AT_END 1
#endif
#ifndef VDBRT_MINBLOCKS_DENSE
#define VDBRT_MINBLOCKS_DENSE 6
#endif
#ifndef VDBRT_MINBLOCKS_LONG
#define VDBRT_MINBLOCKS_LONG 4
#endif
// How a warp is fed (all run-time, warp-uniform; launchLevelSet fills it in):
//   * the queue hands out STRIPS of `strip_tiles` consecutive 8x4 tiles (consecutive tiles of a macro-tile row are neighbours in x);
//   * a warp feeds its lanes from its own strip: when at least `refill` lanes have finished their pixel they get the next
//     pixels of the strip, so lanes whose rays ended early do not idle until the slowest ray of a fixed tile is done -- and
//     the new rays start right next to the ones still running (same leaves, same tree path).  Round 1 measured per-lane
//     refill from the GLOBAL queue as slower: there the next tickets belong to whatever tile the other 2 367 warps have
//     reached, and the lanes of a warp end up in different parts of the tree.  refill = 32: a fresh tile when all lanes are done;
//   * heavy strips first: k_probe_levelset traces one ray per tile with a budget and sorts the strips into two "heavy" lists;
//     the queue's first tickets walk those lists, the rest walk all strips in order and skip the ones already handed out.
//     A silhouette tile (every ray grazes the surface, ~10x the mean tile) that starts when the queue is nearly empty was the
//     tail of the frame: 0.6 of 2.75 ms on C2 with one GPU, half of the frame time of a 1/8 share.
//   * SM-affine queues (`affine` > 0): the strips are dealt out in chunks of `affine` consecutive strips, chunk c to SM c mod #SMs, and
//     every SM drains its own queue (one counter per SM) before it steals from the others: the 16 warps of an SM then work
//     on neighbouring tiles and share their L1 lines (node tables, leaf headers, halo blocks) instead of 16 unrelated tiles.
struct Sched {
    uint32_t strip_tiles, refill, eager;
    uint32_t affine;             // strips per chunk (0: one global queue)
    unsigned int* smq;           // affine: one counter per SM (zeroed before the launch), nq of them
    uint32_t nq;
    uint32_t* cost_out;          // history: SM clock cycles / 16 each tile of THIS frame took (null: not recorded); cost_sum: their sum
    unsigned long long* cost_sum;
    uint32_t* cost_max;          // history: the largest of them (the critical path of the frame's heaviest tile)
    unsigned long long* warp_exit;   // diagnostics (VDBRT_DEBUG_EXIT): %globaltimer of every warp when it leaves the kernel, [0] = launch start
    const uint32_t* ctl;         // [0], [1]: strips in the two heavy lists (null: no ordering)
    const uint32_t* listA; const uint32_t* listB;
    const uint8_t* cls;          // per strip: 0 = not in a list
};
struct OrderBufs { uint32_t* cost; uint32_t* done; uint32_t* ctl; uint32_t* listA; uint32_t* listB; uint8_t* cls; };

__device__ __forceinline__ bool ticketToPixel(const TileMap& m, unsigned ticket, uint32_t& px, uint32_t& py)
{
    const unsigned item = ticket >> 5, slot = ticket & 31u;
    const uint32_t macro = m.rank + (item / m.sub_per_macro) * m.count;
    const uint32_t sub = item % m.sub_per_macro;
    const uint32_t mx = macro % m.macro_x, my = macro / m.macro_x;
    const uint32_t lx = (sub % m.sub_x) * kSubW + (slot % kSubW), ly = (sub / m.sub_x) * kSubH + (slot / kSubW);
    px = mx * m.tile_w + lx; py = my * m.tile_h + ly;
    return lx < m.tile_w && ly < m.tile_h && px < m.width && py < m.height;
}

// getWorldPosAndNml (tools/RayIntersector.h:575-582) + shader + the optional per-pixel records of the primary ray
template<bool AUX>
__device__ __forceinline__ float4 shadeHit(const DevGrid& g, const DevShader& sh, const LsHit& h, const Ray& ray, double wdx, double wdy, double wdz,
                                           size_t pix, const AuxOut& aux, bool record)
{
    // normalise the gradient in double, map the position
    double x = h.px, y = h.py, z = h.pz;
    double nx = h.gx, ny = h.gy, nz = h.gz;
    vnormalize(nx, ny, nz);
    indexToWorldPos(g, x, y, z);
    if (AUX && record) {
        if (aux.ijk) { aux.ijk[3 * pix] = h.ix; aux.ijk[3 * pix + 1] = h.iy; aux.ijk[3 * pix + 2] = h.iz; }
        if (aux.t_index) aux.t_index[pix] = h.time;
        // getWorldTime (:588-591): mTime * |J dir|
        if (aux.t_world) aux.t_world[pix] = h.time * jacobianLength(g, ray.dx, ray.dy, ray.dz);
        if (aux.xyz) { aux.xyz[3 * pix] = x; aux.xyz[3 * pix + 1] = y; aux.xyz[3 * pix + 2] = z; }
        if (aux.nml) { aux.nml[3 * pix] = nx; aux.nml[3 * pix + 1] = ny; aux.nml[3 * pix + 2] = nz; }
    }
    return shade(sh, x, y, z, nx, ny, nz, wdx, wdy, wdz);
}

// ------------------------------------------------------------------------------------------------------------
// Long rays.  A ray that grazes the surface marches hundreds of narrow-band voxels; one such ray per warp keeps the
// whole warp (and, at the end of the frame, the whole GPU) waiting.  The render kernel therefore gives every 8x4 tile a
// BUDGET of warp iterations; rays still running after that are suspended into LongRay records and finished by
// "rounds" of two homogeneous kernels (default: one round of K = 64 leaf visits per ray):
//   scout   one thread per long ray first RESOLVES the previous round -- the first segment with a hit wins -> shade + film;
//           walked out of the grid -> miss -- and otherwise walks the node levels only (root/upper/lower DDAs) and writes
//           the next K leaf visits (time range + leaf handle) it finds as segments;
//   march   one thread per SEGMENT runs the voxel DDA / zero-crossing search of that leaf (LevelSetHDDA<Tree,-1>::test).
// k_long_finish resolves the last round and walks whatever is still alive to its end in line.
// A leaf visit depends only on the ray and its [t0,t1] (the tester is re-initialised per leaf, math/DDA.h:172-173), so
// marching the leaves of one ray in parallel and taking the first hit in visit order is exactly the reference's result.
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kNoHit = 0xffffffffu;
constexpr int kMaxRounds = 8;
constexpr uint32_t kDefaultTail = 24;     // tail rule: warp iterations a tile may still spend once the queue has run dry
constexpr uint32_t kDefaultBudget = 160;  // per-tile rule (VDBRT_LS_TAIL=0): warp iterations per 8x4 tile before its running rays are suspended ...
constexpr uint32_t kDefaultFactor = 50;   // ... or this many percent of a warp's fair share of the launch
constexpr int kDefaultRounds = 1;
constexpr uint32_t kDenseMinTilesPerSm = 800;   // default of ls_dense: without tile costs of a previous frame, the 6-CTA instantiation from this many 8x4 tiles per SM on
constexpr uint32_t kDenseFactor = 100;          // default of ls_dense_factor (percent): with them, when a warp's share of the summed tile times is at least this much of the heaviest tile
constexpr double kRoundsMaxTilesPerSm = 192.0;   // rounds are on by default only below this (12 tiles per warp at 16 warps / SM; see launchLevelSet)

struct LongRay {
    Ray ray;                          // index space, clipped
    double wdx, wdy, wdz;             // world direction (shader input)
    Dda cur; int lvl;                 // the DDA of the level the walk is on (<= 2 while suspended)
    DdaSave park[2];                  // suspended parents: root level, upper level
    int kx, ky, kz; uint32_t n2, n1, n0;
    uint32_t pix, flags;              // LsWalk::kSkip | kStep
    uint32_t segBase, segCount, best, state;   // this round's segments; index of the first one that hit; 1 = walked out of the grid
};
struct SegIn { double t0, t1; int kx, ky, kz; uint32_t n2, n1, n0; };
struct SegOut { double time, px, py, pz; float gx, gy, gz; int ix, iy, iz; };
struct LongCtl { uint32_t nLong, tiles; unsigned long long spent; uint32_t segCount[kMaxRounds]; uint32_t live[kMaxRounds + 1]; };   // live[r]: rays that walked in round r; spent/tiles: iterations of the finished tiles
struct LongBufs { LongRay* rays; uint32_t* liveA; uint32_t* liveB; SegIn* segIn; SegOut* segOut; LongCtl* ctl; uint32_t capLong, capSeg, budget, factor;
                  uint32_t tail, voxel_only; };   // tail > 0: suspend only once the work queue has run dry, `tail` warp iterations after a warp has seen that

// Writes the record of a suspended ray.  Nothing of the caller's state is modified (the record is built from copies), so
// the cold suspension path adds no merges to the registers the render loop carries from iteration to iteration.
template<int THREADS>
__device__ __forceinline__ void suspendRay(LongRay& r, const Ray& ray, double wdx, double wdy, double wdz, const LsWalk& w, const WalkSmem<THREADS>& sm,
                                           const TreeCursor& acc, size_t pix)
{
    const int t = threadIdx.x;
    r.ray = ray; r.wdx = wdx; r.wdy = wdy; r.wdz = wdz;
    uint32_t flags = w.f & (LsWalk::kSkip | LsWalk::kStep);
    if (w.lvl == 3) {
        // rewind to the lower node's DDA standing on this leaf: the scout finds the leaf again and the march redoes the visit
        r.lvl = 2; flags = 0u;
        r.cur.t0 = sm.t0[2][t]; r.cur.t1 = sm.t1[2][t]; r.cur.nx = sm.nx[2][t]; r.cur.ny = sm.ny[2][t]; r.cur.nz = sm.nz[2][t];
        r.cur.vx = sm.vx[2][t]; r.cur.vy = sm.vy[2][t]; r.cur.vz = sm.vz[2][t];
    } else { r.cur = w.cur; r.lvl = w.lvl; }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        r.park[l].t1 = sm.t1[l][t]; r.park[l].nx = sm.nx[l][t]; r.park[l].ny = sm.ny[l][t]; r.park[l].nz = sm.nz[l][t];
        r.park[l].vx = sm.vx[l][t]; r.park[l].vy = sm.vy[l][t]; r.park[l].vz = sm.vz[l][t];
    }
    r.kx = acc.kx; r.ky = acc.ky; r.kz = acc.kz; r.n2 = acc.n2; r.n1 = acc.n1; r.n0 = acc.n0;
    r.pix = uint32_t(pix); r.flags = flags;
    r.segBase = 0u; r.segCount = 0u; r.best = kNoHit; r.state = 0u;
}

template<int THREADS>
__device__ __forceinline__ void resumeRay(const LongRay& r, Ray& ray, LsWalk& w, WalkSmem<THREADS>& sm, TreeCursor& acc)
{
    ray = r.ray;
    w.reset();
    w.cur = r.cur; w.setLevel(r.lvl);
    const int t = threadIdx.x;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        sm.t1[l][t] = r.park[l].t1; sm.nx[l][t] = r.park[l].nx; sm.ny[l][t] = r.park[l].ny; sm.nz[l][t] = r.park[l].nz;
        sm.vx[l][t] = r.park[l].vx; sm.vy[l][t] = r.park[l].vy; sm.vz[l][t] = r.park[l].vz;
    }
    acc.kx = r.kx; acc.ky = r.ky; acc.kz = r.kz; acc.n2 = r.n2; acc.n1 = r.n1; acc.n0 = r.n0;
    w.f = r.flags;
}

// MULTI = false: one sample per pixel (no sample accumulator, sample counter or jitter index in registers).
// REFINE = true: LinearSearchImpl's secant refinements (p.iters > 0), see lsAdvance.
// DENSE = true: 6 CTAs per SM (80 registers, some spills inside the loop) -- for launches with very many tiles per SM, which are bound by
// throughput (C4, 1 751 tiles per SM: 26.4 -> 24.7 ms); with fewer tiles the longer critical path of the heaviest tile costs more than the
// extra warps bring (C2, 438 tiles per SM: 1.76 -> 1.88 ms), see launchLevelSet.
template<bool AUX, bool COUNT, bool LONG, bool MULTI, bool REFINE = false, int LEAF = kLeafFloat, bool DENSE = false>
__global__ void __launch_bounds__(kBlockThreads, LONG ? VDBRT_MINBLOCKS_LONG : (DENSE ? VDBRT_MINBLOCKS_DENSE : VDBRT_MINBLOCKS))
k_render_levelset(const __grid_constant__ DevGrid g, const __grid_constant__ DevCamera cam, const __grid_constant__ DevShader sh,
                  const __grid_constant__ LsParams p, const __grid_constant__ TileMap tm, float4* film,
                  AuxOut aux, unsigned int* queue, unsigned long long* counters, const __grid_constant__ LongBufs lb,
                  const __grid_constant__ Sched sc)
{
    __shared__ RootSmem root;
    __shared__ WalkSmem<kBlockThreads> wsm;
    stageRoot(g, root);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    if (sc.warp_exit && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        sc.warp_exit[0] = now;
    }
    const unsigned total = tm.items * 32u;            // pixel slots: 32 per warp tile
    const unsigned stripSlots = sc.strip_tiles * 32u;
    const unsigned nStrips = (tm.items + sc.strip_tiles - 1u) / sc.strip_tiles;
    const unsigned nA = sc.ctl ? sc.ctl[0] : 0u, nB = sc.ctl ? sc.ctl[1] : 0u;      // written by k_probe_levelset before this launch
    const unsigned nTickets = nA + nB + nStrips;
    const unsigned thr = LONG ? 32u : sc.refill;      // the tile budget of the long-ray rounds counts whole tiles
    unsigned mySm = 0;
    asm("mov.u32 %0, %%smid;" : "=r"(mySm));
    mySm %= (sc.nq ? sc.nq : 1u);
    TreeCursor acc; acc.reset();
    Stencil st; st.reset();
    Counters c = {};
    // per-lane pixel / ray state
    bool hasPix = false, drained = false;               // a lane has a ray while !(walk.f & LsWalk::kIdle)
    uint32_t px = 0, py = 0, k = 0;                      // (the pixel's index in the film is rebuilt from these where it is needed)
    unsigned long long n = 0;
    float4 col = make_float4(0.f, 0.f, 0.f, 1.f);       // the pixel's background is re-read when a ray misses (the film is written once, at the end)
    Ray ray;
    LsWalk walk; LsHit h;
    ray.ex = ray.ey = ray.ez = 0.0; ray.setDir(1.0, 1.0, 1.0); ray.t0 = ray.t1 = 0.0; walk.begin(ray); walk.f = LsWalk::kIdle;
    h.time = 0.0; h.ix = h.iy = h.iz = 0; h.px = h.py = h.pz = 0.0; h.gx = h.gy = h.gz = 0.f;
    walk.cur.t0 = walk.cur.t1 = walk.cur.nx = walk.cur.ny = walk.cur.nz = 0.0; walk.cur.vx = walk.cur.vy = walk.cur.vz = 0;
    unsigned sNext = 0, sEnd = 0;                       // pixel slots of the warp's strip that are still to be handed out (warp-uniform)
    unsigned curStrip = 0xffffffffu, tileT0 = 0;        // history: the strip the warp is working on and the SM clock when it took it (nothing is counted per iteration)

    long long tileStart = 0; unsigned long long tileIters = 0, tileActive = 0;
    // warp iterations spent on the current tile, and what a tile may spend: at least lb.budget, and lb.factor percent of
    // a warp's fair share of the whole launch (tiles per warp x the mean of the tiles the grid has finished so far, two
    // atomics per tile).  Suspending pays when ONE tile is long against everything else a warp has to do -- a small
    // partition of a frame; a launch with plenty of tiles per warp balances itself and suspends next to nothing.
    // Tail rule (lb.tail > 0, the default): nothing is suspended while the queue still has tiles -- a heavy tile that starts early
    // overlaps with everything else and costs nothing.  Once the queue has run dry (every warp polls the ticket counter every 16
    // iterations) the rest of the launch is only as fast as its slowest ray: rays still running `tail` iterations later are
    // suspended and finished by the rounds, which spread one ray's leaf visits over the idle machine.  Measured on the B200
    // (VDBRT_DEBUG_EXIT): the last warp of the C2 frame leaves 0.66 ms after the first (24 % of the frame), of a 1/8 share of
    // the C4 frame 1.2 ms after the first (24 %).
    const float tilesPerWarp = float(tm.items) / float(gridDim.x * (kBlockThreads / 32));
    uint32_t spent = 0, limit = lb.tail ? lb.tail : lb.budget;
    bool tail = false;                  // the queue has run dry
    bool longFull = false;              // no room left for suspended rays: finish everything in line
    // Two nested loops.  The OUTER one runs once per refill: it feeds idle lanes the next pixels of the warp's strip and sets
    // up the next ray of every lane that needs one (the only place the ray registers are written).  The INNER one advances
    // the running rays and finishes the ones that end, until enough lanes are idle (or one wants its next sample) or no lane
    // is running: nothing of the outer loop's bookkeeping is executed per traversal step.
    // shade / composite the ray that ended with `status`, then next sample or write the pixel
    auto finishRay = [&](int& status) {
        const size_t pix = size_t(py) * tm.width + px;
        const bool hit = status == kWalkHit;
        float4 s;
        if (hit) {
            if (COUNT) ++c.hits;
            lsFinishHit<COUNT, REFINE, LEAF>(g, root, acc, st, ray, p.iso, walk, h, c, int(p.iters));
            // the shader wants the WORLD direction of the ray: rebuilt from the camera here (the same arithmetic on the same
            // inputs) rather than carried through the traversal in six registers
            Ray wr;
            const bool first = !MULTI || k == 0;
            cameraRay(cam, px, py, first ? 0.5 : p.jitter[(n - 2) & 15], first ? 0.5 : p.jitter[(n - 1) & 15], wr);
            s = shadeHit<AUX>(g, sh, h, ray, wr.dx, wr.dy, wr.dz, pix, aux, first);
        } else s = p.uniform_bg ? make_float4(p.bg[0], p.bg[1], p.bg[2], p.bg[3]) : p.bg_film[pix];
        if (AUX && (!MULTI || k == 0) && aux.hit) aux.hit[pix] = hit ? 1 : 0;
        if (MULTI) {
            if (k == 0) col = s;
            else { col.x += s.x; col.y += s.y; col.z += s.z; col.w += s.w; }     // RGBA::operator+= (:250)
            if (++k > p.sub) {
                film[pix] = make_float4(col.x * p.frac, col.y * p.frac, col.z * p.frac, 1.0f);   // bg = c*frac, alpha rebuilt as 1 (:247)
                hasPix = false;
            }
        } else {
            film[pix] = make_float4(s.x * p.frac, s.y * p.frac, s.z * p.frac, 1.0f);
            hasPix = false;
        }
        status = kWalkContinue;
    };
    for (;;) {
        __syncwarp();
        // (0) the tile has used up its budget (the inner loop leaves when that happens): suspend the rays that are still running,
        // they continue in the long-ray rounds.  Kept out of the inner loop so that the hot loop is the same with and without it.
        if (LONG && spent > limit && !longFull && (!lb.tail || tail)) {
            // tail rule with lb.voxel_only: only rays that are marching voxels go
            // to the rounds (the rounds parallelise leaf marches; a ray that is crossing empty nodes is walked by the scout no faster)
            const bool sus = !(walk.f & LsWalk::kIdle) && (!(lb.tail && lb.voxel_only) || walk.lvl == 3);
            const unsigned m = __ballot_sync(0xffffffffu, sus);
            if (lb.tail) spent = 0;                                 // the lanes that stay are looked at again `tail` iterations later
            if (m && sc.cost_out && lane == 0 && curStrip != 0xffffffffu) { sc.cost_out[curStrip] = 0x7fffffffu; curStrip = 0xffffffffu; }
            if (m) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&lb.ctl->nLong, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned idx = base + __popc(m & ((1u << lane) - 1u));
                if (base + __popc(m) > lb.capLong) longFull = true;
                if (sus && idx < lb.capLong) {
                    Ray wr;
                    cameraRay(cam, px, py, 0.5, 0.5, wr);                 // LONG kernels are one sample per pixel
                    suspendRay(lb.rays[idx], ray, wr.dx, wr.dy, wr.dz, walk, wsm, acc, size_t(py) * tm.width + px);
                    walk.f = LsWalk::kIdle; hasPix = false;
                }
            }
        }
        // (1) feed idle lanes
        const unsigned idle = __ballot_sync(0xffffffffu, !hasPix);
        const unsigned nIdle = __popc(idle);
        if (COUNT && idle == 0xffffffffu && lane == 0) {
            const long long now = clock64();
            if (tileStart) {
                atomicMax(counters + 10, (unsigned long long)(now - tileStart)); atomicMax(counters + 11, tileIters); atomicAdd(counters + 12, (unsigned long long)(now - tileStart)); atomicAdd(counters + 13, 1ull);
                if (now - tileStart > 1500000) { atomicAdd(counters + 14, 1ull); atomicAdd(counters + 15, tileActive * 100ull / (tileIters ? tileIters : 1)); }
            }
            tileStart = now; tileIters = 0; tileActive = 0;
        }
        if (sc.cost_out && idle == 0xffffffffu && sNext >= sEnd && curStrip != 0xffffffffu) {
            if (lane == 0) {
                const unsigned cost = (unsigned(clock()) - tileT0) >> 4;            // 16-cycle units: stays far below the 'suspended' mark
                sc.cost_out[curStrip] = cost; atomicAdd(sc.cost_sum, (unsigned long long)cost); atomicMax(sc.cost_max, cost);
            }
            curStrip = 0xffffffffu;
        }
        if (idle == 0xffffffffu && drained && sNext >= sEnd) break;
        if (LONG && idle == 0xffffffffu && !lb.tail) {
            unsigned m = 0;
            if (lane == 0) {
                if (spent > 1u) { atomicAdd(&lb.ctl->spent, (unsigned long long)(spent < limit ? spent : limit)); atomicAdd(&lb.ctl->tiles, 1u); }
                const unsigned long long sum = *reinterpret_cast<volatile unsigned long long*>(&lb.ctl->spent);
                const uint32_t nt = *reinterpret_cast<volatile uint32_t*>(&lb.ctl->tiles);
                m = nt >= 64u ? uint32_t(0.01f * float(lb.factor) * tilesPerWarp * float(sum) / float(nt)) : 0u;
            }
            m = __shfl_sync(0xffffffffu, m, 0);
            limit = m > lb.budget ? m : lb.budget;
            spent = 0;
        }
        // the warp's strip is used up: the next one comes from the queue -- when every lane is idle, or (eager) as soon as a refill is due
        if (sNext >= sEnd && !drained && (idle == 0xffffffffu || (sc.eager && nIdle >= thr))) {
            unsigned s = 0xffffffffu;
            if (lane == 0 && sc.affine) {
                // own queue first, then the other SMs' (a queue that is used up is skipped after one plain load)
                const unsigned chunk = sc.affine, perRound = chunk * sc.nq;
                for (unsigned v = 0; v < sc.nq && s == 0xffffffffu; ++v) {
                    const unsigned q = (mySm + v) % sc.nq;
                    for (;;) {
                        // position p of queue q is strip (p / chunk) * perRound + q * chunk + p % chunk; the queue is used up when its chunk starts past the end
                        const unsigned seen = *reinterpret_cast<volatile unsigned int*>(sc.smq + q);
                        if ((seen / chunk) * perRound + q * chunk >= nStrips) break;
                        const unsigned pos = atomicAdd(sc.smq + q, 1u);
                        const unsigned start = (pos / chunk) * perRound + q * chunk;
                        if (start >= nStrips) break;
                        if (start + pos % chunk < nStrips) { s = start + pos % chunk; break; }
                    }
                }
            } else if (lane == 0) {
                for (;;) {
                    const unsigned t = atomicAdd(queue, 1u);
                    if (t >= nTickets) break;
                    if (t < nA) { s = sc.listA[t]; break; }                   // heavy strips first (k_probe_levelset)
                    if (t < nA + nB) { s = sc.listB[t - nA]; break; }
                    const unsigned u = t - nA - nB;
                    if (!sc.cls || !sc.cls[u]) { s = u; break; }              // everything else in tile order
                }
            }
            s = __shfl_sync(0xffffffffu, s, 0);
            if (s == 0xffffffffu) { drained = true; tail = true; }
            else { sNext = s * stripSlots; sEnd = sNext + stripSlots < total ? sNext + stripSlots : total; curStrip = s; tileT0 = unsigned(clock()); }
        }
        if (sNext < sEnd && (idle == 0xffffffffu || nIdle >= thr)) {
            if (!hasPix) {
                const unsigned ticket = sNext + __popc(idle & ((1u << lane) - 1u));
                if (ticket < sEnd && ticketToPixel(tm, ticket, px, py)) {
                    hasPix = true; walk.f = LsWalk::kIdle; k = 0;
                    if (MULTI) n = 2ull * p.sub * (size_t(py) * tm.width + px);
                }
            }
            sNext += nIdle;
            if (!__any_sync(0xffffffffu, hasPix)) continue;                    // slots outside the film (edge tiles)
        }
        // (2) start the next ray of the lane's pixel
        int status = kWalkContinue;
        if (hasPix && (walk.f & LsWalk::kIdle)) {
            const bool first = !MULTI || k == 0;
            cameraRay(cam, px, py, first ? 0.5 : p.jitter[n & 15], first ? 0.5 : p.jitter[(n + 1) & 15], ray);
            if (MULTI && !first) n += 2;
            if (COUNT) ++c.rays;
            // intersectsWS: setWorldRay = worldToIndex + clip (tools/RayIntersector.h:558-562)
            worldToIndex(g, ray);
            if (clipRay(ray, g, 0)) walk.begin(ray);                 // clears kIdle
            else status = kWalkMiss;
        }
        const bool canRefill = thr < 32u && (sNext < sEnd || (sc.eager && !drained));
#pragma unroll 1
        for (;;) {
            __syncwarp();
            if (COUNT) {
                const unsigned live = __popc(__ballot_sync(0xffffffffu, !(walk.f & LsWalk::kIdle)));
                ++tileIters; tileActive += live;
                if (lane == 0) atomicAdd(counters + 16 + (live + 3u) / 4u, 1ull);      // histogram of running lanes per iteration: 0, 1-4, 5-8, ..., 29-32
            }
            if (LONG) {
                ++spent;
                if (lb.tail && !tail && (spent & 15u) == 0u) {
                    tail = *reinterpret_cast<volatile unsigned int*>(queue) >= nTickets;
                    if (tail) spent = 0;
                }
            }
            // (3) advance running rays by one step (all lanes call it: it re-synchronises the warp between its phases).
            // Deferring the rare phases (level set-up, stencil) until several lanes want them was measured: no gain.
            {
                const int r = lsAdvance<COUNT, true, kBlockThreads, REFINE, LEAF>(g, root, wsm, acc, st, ray, p.iso, p.vmin, p.vmax, walk, h, c, int(p.iters));
                if (r != kWalkContinue) status = r;          // (idle lanes return kWalkContinue and keep the status of their finished ray)
            }
            __syncwarp();
            // (4) a ray ended.  One sample per pixel: the lane just stops; its pixel is shaded and written after the loop, together with
            // the other pixels of the tile (VDBRT_SHADE_AT_END) -- otherwise: shade / composite now, then next sample or write the pixel
            if (status != kWalkContinue) {
                walk.f = LsWalk::kIdle;
                if (MULTI || !VDBRT_SHADE_AT_END) finishRay(status);
            }
            // back to the outer loop when a lane wants its next ray, when enough lanes are idle and the strip has pixels for them,
            // or when nothing is running any more
            const unsigned running = __ballot_sync(0xffffffffu, !(walk.f & LsWalk::kIdle));
            if (running == 0u || (canRefill && 32u - (unsigned)__popc(running) >= thr) || (MULTI && __any_sync(0xffffffffu, hasPix && (walk.f & LsWalk::kIdle)))) break;
            if (LONG && spent > limit && !longFull && (!lb.tail || tail)) break;
        }
        if (!MULTI && VDBRT_SHADE_AT_END) {
            __syncwarp();
            if (hasPix && status != kWalkContinue) finishRay(status);
        }
    }
    if (COUNT) { flushCounters(c, counters); flushDiag(c, counters + 32); }
    if (sc.warp_exit && lane == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        sc.warp_exit[1 + blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5)] = now;
    }
}

// Heavy tiles first from HISTORY: the time (SM clock) every tile took in the previous frame this context rendered with the same
// film, tiles and partition (Sched::cost_out).  Tiles that took at least fA / fB times the mean go to list A / B; the render
// kernel hands those out first (Sched::ctl, listA, listB, cls) and everything else in tile order.  Costs nothing but this launch
// (one thread per tile) and needs no probe rays; the first frame of a sequence is rendered in plain tile order.
__global__ void k_order_from_history(const uint32_t* __restrict__ cost, const unsigned long long* __restrict__ sum, uint32_t items, const __grid_constant__ OrderBufs ob,
                                     float fA, float fB)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= items) return;
    const float mean = float(*sum) / float(items);
    const float c = float(cost[i]);
    if (c >= fA * mean && c > 8.f) { ob.cls[i] = 1; ob.listA[atomicAdd(ob.ctl + 0, 1u)] = i; }
    else if (c >= fB * mean && c > 8.f) { ob.cls[i] = 2; ob.listB[atomicAdd(ob.ctl + 1, 1u)] = i; }
}

// One ray per 8x4 tile (the pixel in its middle), traced for at most `cap` steps: the number of steps is the tile's cost
// estimate.  The last tile of a strip to finish sorts the strip into list A (cost >= thrA) or B (>= thrB) -- see Sched.
// Plain per-thread loop (a launch of tiles/32 warps, ~3 % of the frame's rays); nothing it computes reaches the film.
__global__ void __launch_bounds__(kBlockThreads)
k_probe_levelset(const __grid_constant__ DevGrid g, const __grid_constant__ DevCamera cam, const __grid_constant__ LsParams p,
                 const __grid_constant__ TileMap tm, const __grid_constant__ OrderBufs ob, uint32_t stripTiles, uint32_t cap, uint32_t thrA, uint32_t thrB)
{
    __shared__ RootSmem root;
    __shared__ WalkSmem<kBlockThreads> wsm;
    stageRoot(g, root);
    __syncthreads();
    Counters c = {};
    for (uint32_t item = blockIdx.x * kBlockThreads + threadIdx.x; item < tm.items; item += gridDim.x * kBlockThreads) {
        uint32_t px, py, it = 0;
        if (ticketToPixel(tm, item * 32u + 20u, px, py)) {
            Ray ray;
            cameraRay(cam, px, py, 0.5, 0.5, ray);
            worldToIndex(g, ray);
            if (clipRay(ray, g, 0)) {
                TreeCursor acc; acc.reset();
                Stencil st; st.reset();
                LsWalk w; w.begin(ray);
                LsHit h = {};
#pragma unroll 1
                for (; it < cap; ++it)
                    if (lsAdvance<false, false, kBlockThreads>(g, root, wsm, acc, st, ray, p.iso, p.vmin, p.vmax, w, h, c) != kWalkContinue) break;
            }
        }
        const uint32_t strip = item / stripTiles;
        const uint32_t first = strip * stripTiles, nTiles = tm.items - first < stripTiles ? tm.items - first : stripTiles;
        atomicMax(ob.cost + strip, it);
        __threadfence();
        if (atomicAdd(ob.done + strip, 1u) == nTiles - 1u) {
            const uint32_t cost = atomicMax(ob.cost + strip, 0u);
            if (cost >= thrA) { ob.cls[strip] = 1; ob.listA[atomicAdd(ob.ctl + 0, 1u)] = strip; }
            else if (cost >= thrB) { ob.cls[strip] = 2; ob.listB[atomicAdd(ob.ctl + 1, 1u)] = strip; }
        }
    }
}

// Round r works on the rays that walked in round r-1 (list r-1; all suspended rays for r = 0).  Its scout first RESOLVES
// round r-1 for its ray -- the first segment that hit wins, a ray that left the grid without a hit is a miss -- and only
// an unresolved ray walks on; those rays form list r, which the march kernel of round r and the scout of round r+1 read.
__device__ __forceinline__ uint32_t liveCount(const LongBufs& lb, int round)
{
    const uint32_t n = round == 0 ? lb.ctl->nLong : lb.ctl->live[round - 1];
    return n < lb.capLong ? n : lb.capLong;
}
__device__ __forceinline__ const uint32_t* liveList(const LongBufs& lb, int round) { return round == 0 ? nullptr : (((round - 1) & 1) ? lb.liveB : lb.liveA); }

template<bool AUX>
__device__ __forceinline__ void writeLongPixel(const DevGrid& g, const DevShader& sh, const LsParams& p, float4* film, const AuxOut& aux, const LongRay& r,
                                               bool hit, const LsHit& h)
{
    const size_t pix = r.pix;
    float4 s = p.uniform_bg ? make_float4(p.bg[0], p.bg[1], p.bg[2], p.bg[3]) : p.bg_film[pix];
    if (hit) s = shadeHit<AUX>(g, sh, h, r.ray, r.wdx, r.wdy, r.wdz, pix, aux, true);
    if (AUX && aux.hit) aux.hit[pix] = hit ? 1 : 0;
    film[pix] = make_float4(s.x * p.frac, s.y * p.frac, s.z * p.frac, 1.0f);      // one sample per pixel: c = s, bg = c*frac, alpha 1
}

// true when the ray is finished (pixel written)
template<bool AUX>
__device__ __forceinline__ bool resolveLong(const DevGrid& g, const DevShader& sh, const LsParams& p, float4* film, const AuxOut& aux, const LongBufs& lb, LongRay* r)
{
    if (r->best != kNoHit) {
        const SegOut o = lb.segOut[r->segBase + r->best];
        LsHit h;
        h.time = o.time; h.ix = o.ix; h.iy = o.iy; h.iz = o.iz; h.px = o.px; h.py = o.py; h.pz = o.pz; h.gx = o.gx; h.gy = o.gy; h.gz = o.gz;
        writeLongPixel<AUX>(g, sh, p, film, aux, *r, true, h);
        return true;
    }
    if (r->state == 1u) {
        LsHit h = {};
        writeLongPixel<AUX>(g, sh, p, film, aux, *r, false, h);
        return true;
    }
    return false;
}

// The scout is a plain per-thread loop (no phases, no warp synchronisation): LevelSetHDDA<Tree,2/1/0>::test (math/DDA.h:144-165)
// with "enter the leaf" replaced by "write a segment".  Its state between rounds is "about to probe the current cell of
// the DDA at `lvl`" plus what the render kernel may have left pending (a level to initialise, a cell to step past).
// The walk of one ray is a sequential chain, and 32 different walks in one warp run one after the other; when there
// are fewer rays than lanes in the grid the rays are therefore spread out, down to one ray per warp.
template<bool AUX>
__global__ void __launch_bounds__(kBlockThreads)
k_long_scout(const __grid_constant__ DevGrid g, const __grid_constant__ DevShader sh, const __grid_constant__ LsParams p, float4* __restrict__ film,
             AuxOut aux, const __grid_constant__ LongBufs lb, int round, uint32_t K)
{
    __shared__ RootSmem root;
    __shared__ WalkSmem<kBlockThreads> wsm;
    stageRoot(g, root);
    __syncthreads();
    const uint32_t nLive = liveCount(lb, round);
    const uint32_t* list = liveList(lb, round);
    uint32_t* next = (round & 1) ? lb.liveB : lb.liveA;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = gridDim.x * (kBlockThreads / 32);
    uint32_t rpw = (nLive + warps - 1) / warps;                        // rays per warp
    rpw = rpw < 1u ? 1u : (rpw > 32u ? 32u : rpw);
    for (uint32_t base = (blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5)) * rpw; base < nLive; base += warps * rpw) {
        const uint32_t i = base + lane;
        if (lane >= rpw || i >= nLive) continue;
        const uint32_t idx = list ? list[i] : i;
        LongRay* r = lb.rays + idx;
        if (round > 0 && resolveLong<AUX>(g, sh, p, film, aux, lb, r)) continue;
        next[atomicAdd(&lb.ctl->live[round], 1u)] = idx;
        const uint32_t sb = atomicAdd(&lb.ctl->segCount[round], K);    // K consecutive segment slots
        if (sb + K > lb.capSeg) { r->segCount = 0u; r->best = kNoHit; continue; }      // no room this round: try again in the next one
        Ray ray; LsWalk walk; TreeCursor acc;
        resumeRay(*r, ray, walk, wsm, acc);
        Dda& cur = walk.cur;
        uint32_t emitted = 0;
        bool exhausted = false;
        bool advance = (walk.f & (LsWalk::kStep | LsWalk::kSkip)) != 0u;
#pragma unroll 1
        for (;;) {
            if (advance) {
                // while (dda.step()); an exhausted level returns to its parent, which steps in turn (DDA.h:158-159)
                bool alive = true, failed = !cur.step(ray, LsWalk::shiftOf(walk.lvl));
#pragma unroll 1
                for (;;) {
#pragma unroll 1
                    while (failed) {
                        if (walk.lvl == 0) { alive = false; break; }
                        --walk.lvl; wsm.unpark(walk.lvl, cur);
                        failed = !cur.step(ray, LsWalk::shiftOf(walk.lvl));
                    }
                    if (!alive || emitted == K || walk.lvl != 2 || !acc.n1) break;
                    // short cut: run through the empty cells of the lower node the cursor is in (one child-mask word per cell);
                    // a cell with a leaf, or one that rounding pushed outside the node, goes to the general probe below
                    const uint8_t* cm = TreeCursor::node(g, acc.n1) + kLowerCMask;
#pragma unroll 1
                    for (;;) {
                        if (uint32_t((cur.vx ^ acc.kx) | (cur.vy ^ acc.ky) | (cur.vz ^ acc.kz)) & ~127u) break;
                        const uint32_t n = lowerOffset(cur.vx, cur.vy, cur.vz);
                        if ((ldg64(cm + 8u * (n >> 6)) >> (n & 63u)) & 1ull) break;
                        if (!cur.step(ray, 3)) { failed = true; break; }
                    }
                    if (!failed) break;
                }
                if (!alive) { exhausted = true; break; }
                if (emitted == K) break;
            }
            advance = true;
            // tester.hasNode<NodeT>(dda.voxel()) (tools/RayIntersector.h:609-613)
            const int depth = acc.descend(g, root, cur.vx, cur.vy, cur.vz);
            if (depth <= 2 - walk.lvl) {
                if (walk.lvl == 2) {
                    SegIn sg;
                    sg.t0 = cur.t0; sg.t1 = cur.next(); sg.kx = acc.kx; sg.ky = acc.ky; sg.kz = acc.kz; sg.n2 = acc.n2; sg.n1 = acc.n1; sg.n0 = acc.n0;
                    lb.segIn[sb + emitted] = sg;
                    ++emitted;                                   // the DDA steps past the leaf before the state is saved
                } else {
                    const double c0 = cur.t0, c1 = cur.next();    // tester.setRange(dda.time(), dda.next()) (DDA.h:154)
                    wsm.park(walk.lvl, cur);
                    ++walk.lvl;
                    cur.init(ray, c0, c1, LsWalk::shiftOf(walk.lvl));
                    advance = false;
                }
            }
        }
        walk.f = 0u;
        suspendRay(*r, ray, r->wdx, r->wdy, r->wdz, walk, wsm, acc, size_t(r->pix));
        r->segBase = sb; r->segCount = emitted; r->state = exhausted ? 1u : 0u;
    }
}

// LevelSetHDDA<TreeT,-1>::test (math/DDA.h:166-177) with LinearSearchImpl (tools/RayIntersector.h:597-657) on one leaf visit
__global__ void __launch_bounds__(kBlockThreads)
k_long_march(const __grid_constant__ DevGrid g, const __grid_constant__ LongBufs lb, int round, uint32_t K, float iso, float vmin, float vmax)
{
    __shared__ RootSmem root;
    stageRoot(g, root);
    __syncthreads();
    const uint32_t nLive = liveCount(lb, round + 1);                  // the rays that walked in this round
    const uint32_t* list = liveList(lb, round + 1);
    const unsigned long long total = (unsigned long long)nLive * K;
    Counters c = {};
    for (unsigned long long t = blockIdx.x * (unsigned long long)kBlockThreads + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * kBlockThreads) {
        const uint32_t i = uint32_t(t / K), j = uint32_t(t % K);
        LongRay* r = lb.rays + list[i];
        if (j >= r->segCount) continue;
        const uint32_t slot = r->segBase + j;
        const SegIn sg = lb.segIn[slot];
        Ray ray = r->ray;
        TreeCursor acc; acc.kx = sg.kx; acc.ky = sg.ky; acc.kz = sg.kz; acc.n2 = sg.n2; acc.n1 = sg.n1; acc.n0 = sg.n0;
        Stencil st; st.reset();
        Dda dda; dda.init(ray, sg.t0, sg.t1, 0);
        // tester.init(dda.time()) (:597-601) is evaluated with the first voxel that passes the value gate (see lsAdvance)
        double T0 = sg.t0, px, py, pz;
        float V0 = 0.f;
        bool lazyInit = true, hit = false;
        SegOut o;
        do {
            // tester(dda.voxel(), dda.next()) (:620-644)
            float V;
            const int depth = acc.descend(g, root, dda.vx, dda.vy, dda.vz);
            if (acc.valueAt(g, root, depth, dda.vx, dda.vy, dda.vz, V) && V > vmin && V < vmax) {
                if (lazyInit) {
                    px = ray.ex + ray.dx * T0; py = ray.ey + ray.dy * T0; pz = ray.ez + ray.dz * T0;
                    st.template moveTo<false>(g, root, acc, px, py, pz, c);
                    V0 = st.interpolation(px, py, pz) - iso;
                    lazyInit = false;
                }
                const double tq = dda.next();
                px = ray.ex + ray.dx * tq; py = ray.ey + ray.dy * tq; pz = ray.ez + ray.dz * tq;
                st.template moveTo<false>(g, root, acc, px, py, pz, c);
                const float V1 = st.interpolation(px, py, pz) - iso;
                if (V0 * V1 <= 0.0f) {
                    o.time = T0 + (tq - T0) * V0 / (V0 - V1);
                    o.ix = dda.vx; o.iy = dda.vy; o.iz = dda.vz;
                    hit = true;
                    break;
                }
                T0 = tq; V0 = V1;
            }
        } while (dda.step(ray, 0));
        if (hit) {
            // getWorldPosAndNml (:575-582): position and stencil gradient at the hit time
            o.px = ray.ex + ray.dx * o.time; o.py = ray.ey + ray.dy * o.time; o.pz = ray.ez + ray.dz * o.time;
            st.template moveTo<false>(g, root, acc, o.px, o.py, o.pz, c);
            st.gradient(g, o.px, o.py, o.pz, o.gx, o.gy, o.gz);
            lb.segOut[slot] = o;
            atomicMin(&r->best, j);
        }
    }
}

// after the last round: resolve it; rays that are still alive are walked to their end in line
template<bool AUX>
__global__ void __launch_bounds__(kBlockThreads)
k_long_finish(const __grid_constant__ DevGrid g, const __grid_constant__ DevShader sh, const __grid_constant__ LsParams p, float4* __restrict__ film,
              AuxOut aux, const __grid_constant__ LongBufs lb, int round)
{
    __shared__ RootSmem root;
    __shared__ WalkSmem<kBlockThreads> wsm;
    stageRoot(g, root);
    __syncthreads();
    const uint32_t nLive = liveCount(lb, round);
    const uint32_t* list = liveList(lb, round);
    Counters c = {};
    for (uint32_t i = blockIdx.x * kBlockThreads + threadIdx.x; i < nLive; i += gridDim.x * kBlockThreads) {
        LongRay* r = lb.rays + (list ? list[i] : i);
        if (round > 0 && resolveLong<AUX>(g, sh, p, film, aux, lb, r)) continue;
        Ray ray; LsWalk walk; TreeCursor acc; Stencil st; LsHit h = {};
        acc.reset(); st.reset();
        resumeRay(*r, ray, walk, wsm, acc);
        int status;
#pragma unroll 1
        do { status = lsAdvance<false, false, kBlockThreads>(g, root, wsm, acc, st, ray, p.iso, p.vmin, p.vmax, walk, h, c); } while (status == kWalkContinue);
        if (status == kWalkHit) lsFinishHit<false, false>(g, root, acc, st, ray, p.iso, walk, h, c);
        writeLongPixel<AUX>(g, sh, p, film, aux, *r, status == kWalkHit, h);
    }
}

// LevelSetRayIntersector::intersectsWS / intersectsIS on arbitrary rays (tools/RayIntersector.h:119-240)
struct HitOut { int32_t hit; int32_t ijk[3]; double t_index, t_world; double xyz_index[3], xyz_world[3], nml[3]; };
struct RayIn { double eye[3], dir[3], t0, t1; };

template<int LEAF = kLeafFloat>
__global__ void __launch_bounds__(kBlockThreads)
k_intersect_levelset(const __grid_constant__ DevGrid g, const RayIn* __restrict__ rays, unsigned long long n, uint32_t space,
                     float iso, float vmin, float vmax, HitOut* __restrict__ hits, int iters)
{
    __shared__ RootSmem root;
    __shared__ WalkSmem<kBlockThreads> wsm;
    stageRoot(g, root);
    __syncthreads();
    TreeCursor acc; acc.reset();
    Stencil st; st.reset();
    Counters c = {};
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
        Ray ray;
        ray.ex = rays[k].eye[0]; ray.ey = rays[k].eye[1]; ray.ez = rays[k].eye[2];
        ray.setDir(rays[k].dir[0], rays[k].dir[1], rays[k].dir[2]);
        ray.t0 = rays[k].t0; ray.t1 = rays[k].t1;
        if (space == 0) worldToIndex(g, ray);
        HitOut o = {};
        LsHit h;
        if (clipRay(ray, g, 0) && intersectLevelSet<false, kBlockThreads, LEAF>(g, root, wsm, acc, st, ray, iso, vmin, vmax, h, c, iters)) {
            double x = h.px, y = h.py, z = h.pz;
            double nx = h.gx, ny = h.gy, nz = h.gz;
            vnormalize(nx, ny, nz);
            o.hit = 1; o.ijk[0] = h.ix; o.ijk[1] = h.iy; o.ijk[2] = h.iz;
            o.t_index = h.time;
            o.t_world = h.time * jacobianLength(g, ray.dx, ray.dy, ray.dz);
            o.xyz_index[0] = x; o.xyz_index[1] = y; o.xyz_index[2] = z;
            indexToWorldPos(g, x, y, z);
            o.xyz_world[0] = x; o.xyz_world[1] = y; o.xyz_world[2] = z;
            o.nml[0] = nx; o.nml[1] = ny; o.nml[2] = nz;
        }
        hits[k] = o;
    }
}

// VolumeRayIntersector::hits on arbitrary rays (tools/RayIntersector.h:368-432)
__global__ void __launch_bounds__(kBlockThreads)
k_volume_spans(const __grid_constant__ DevGrid g, const RayIn* __restrict__ rays, unsigned long long n, uint32_t space,
               uint32_t maxSpans, double* __restrict__ spans, int32_t* __restrict__ counts)
{
    __shared__ RootSmem root;
    __shared__ FogSmem<kBlockThreads> fsm;
    stageRoot(g, root);
    __syncthreads();
    TreeCursor acc; acc.reset();
    Counters c = {};
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
        Ray ray;
        ray.ex = rays[k].eye[0]; ray.ey = rays[k].eye[1]; ray.ez = rays[k].eye[2];
        ray.setDir(rays[k].dir[0], rays[k].dir[1], rays[k].dir[2]);
        ray.t0 = rays[k].t0; ray.t1 = rays[k].t1;
        if (space == 0) worldToIndex(g, ray);
        if (!clipRay(ray, g, 1)) { counts[k] = -1; continue; }
        SpanWalk w; w.begin(ray);
        int cnt = 0; double a, b;
#pragma unroll 1
        for (;;) {
            const int r = w.template advance<false>(g, root, fsm, 0, acc, ray, a, b, c);
            if (r == kSpanDone) break;
            if (r == kSpanEmit) {
                if (uint32_t(cnt) < maxSpans) { spans[(k * maxSpans + cnt) * 2] = a; spans[(k * maxSpans + cnt) * 2 + 1] = b; }
                ++cnt;
            }
        }
        counts[k] = cnt;
    }
}

// ------------------------------------------------------------------------------------------------------------
// VolumeRender<VolumeRayIntersector<FloatGrid>,BoxSampler>::operator() (tools/RayTracer.h:991-1070)
// ------------------------------------------------------------------------------------------------------------
struct VolParams {
    double pstep, sstep, cutoff, gain;
    double light[3];            // world-space unit light direction
    double ext[3], albedo[3];   // extinction = -scattering-absorption; albedo = lightColor*scattering/(scattering+absorption)
    uint32_t sub; float frac;   // samples per pixel - 1, 1/samples (EXTENSION: the reference's VolumeRender has one sample per pixel)
    double jitter[16];
    // the shadow ray's index-space direction, 1/direction and times -- sRay(Vec3R(0), mLightDir) through worldToIndex (tools/RayTracer.h:1017,
    // 1039-1040; Ray ctor defaults t0 = 1e-9, t1 = max, math/Ray.h:57-63) -- the same for every sample, evaluated once on the host with
    // the device's arithmetic (IEEE division and sqrt, no contraction): as kernel parameters they cost the shadow kernel no registers
    double sb[8];
};

// Per-lane state machine, warp-synchronous like the level-set kernel.  A lane works on its primary ray or on a shadow
// ray, and only ONE walker and ONE ray live in its registers: the suspended primary walk / ray sit in shared memory
// while a shadow ray runs.  Each iteration runs the phases  walk | sample | exp  once, each a single site in the code.
//
// Lazy spans: the reference first collects every span of the whole chord (VolumeRayIntersector::hits) and then marches
// them.  Here the walk and the march are interleaved: while a span is still OPEN its end cannot lie before the entry
// time of the last probed cell (`bound`), so every sample time <= bound is certain to be inside the final span and is
// taken right away -- same samples, same order, same arithmetic.  A ray that saturates (|T|^2 < cutoff) after a few
// samples never walks the rest of its chord (for shadow rays inside the fog that is > 90 % of the node probes).
enum { kFogIdle = 0, kFogPrimary = 1, kFogShadow = 2 };
struct VolTiles { uint32_t tile0, tile1; const uint8_t* only; const unsigned int* gate; };   // work items [tile0, tile1); only[t - tile0] != 0 if given; *gate != 0 if given
constexpr int kFogBatch = 8;             // a phase runs when this many lanes want it, or when the other phases are starved

template<bool COUNT, int LEAF = kLeafFloat>
__global__ void __launch_bounds__(kBlockThreads, 3)
k_render_volume(const __grid_constant__ DevGrid g, const __grid_constant__ DevCamera cam, const __grid_constant__ VolParams p,
                const __grid_constant__ TileMap tm, float4* __restrict__ film, unsigned int* queue, unsigned long long* counters,
                const __grid_constant__ VolTiles vt)
{
    // as the fall-back of the wavefront path (vdbrt_fog.cuh): only the flagged tiles of one batch, and nothing at all when none is flagged
    if (vt.gate && *vt.gate == 0u) return;
    __shared__ RootSmem root;
    __shared__ FogSmem<kBlockThreads> sm;
    stageRoot(g, root);
    if (threadIdx.x == 0) {
        // the shadow ray's direction is the same for every sample: sRay(Vec3R(0), mLightDir) through worldToIndex
        // (tools/RayTracer.h:1017,1039-1040; Ray ctor defaults t0 = 1e-9, t1 = max, math/Ray.h:57-63)
        double jx = p.light[0], jy = p.light[1], jz = p.light[2];
        if (g.general) mul3(g.imat, jx, jy, jz); else { jx *= g.inv[0]; jy *= g.inv[1]; jz *= g.inv[2]; }
        const double len = vlength(jx, jy, jz);
        const double dx = jx / len, dy = jy / len, dz = jz / len;
        sm.sbase[0] = dx; sm.sbase[1] = dy; sm.sbase[2] = dz;
        sm.sbase[3] = 1 / dx; sm.sbase[4] = 1 / dy; sm.sbase[5] = 1 / dz;
        sm.sbase[6] = len * 1e-9; sm.sbase[7] = len * DBL_MAX;
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const int tid = threadIdx.x;
    TreeCursor accW, accV;               // walker cursor, sampler cursor (the reference keeps separate accessors too)
    accW.reset(); accV.reset();
    Counters c = {};
    int mode = kFogIdle, pendExp = 0;
    int span = 0;                        // 0: looking for a span, 1: inside an open span (end >= walk.bound), 2: span closed at tend
    bool drained = false;
    bool needRay = false;                // the lane's pixel wants its next sample
    size_t pix = 0;
    uint32_t k = 0;                      // samples of the pixel taken so far
    unsigned long long n = 0;            // jitter index, as in LevelSetRayTracer::operator() (tools/RayTracer.h:903-913)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    Ray ray; SpanWalk walk;
    ray.ex = ray.ey = ray.ez = 0.0; ray.setDir(1.0, 1.0, 1.0); ray.t0 = ray.t1 = 0.0;
    walk.begin(ray); walk.lvl = -1;
    walk.cur.t0 = walk.cur.t1 = walk.cur.nx = walk.cur.ny = walk.cur.nz = 0.0; walk.cur.vx = walk.cur.vy = walk.cur.vz = 0;
    double tcur = 0.0, tend = 0.0;       // march time / end of the closed span of the ACTIVE ray
    double dens = 0.0;                   // density of the sample waiting for its exp
    double Tx = 1.0, Ty = 1.0, Tz = 1.0, Lx = 0.0, Ly = 0.0, Lz = 0.0;      // pTrans, pLumi
    double Sx = 1.0, Sy = 1.0, Sz = 1.0, dTx = 1.0, dTy = 1.0, dTz = 1.0;  // sTrans, dT of the primary sample being lit

    // OUTER loop: tile refill and the set-up of the next sample's primary ray (camera, world->index, clip);
    // INNER loop: walk | sample | exp | luminance | pixel, until a lane wants its next primary ray or no lane is busy.
    for (;;) {
        __syncwarp();
        // (1) a warp takes a fresh 8x4 tile when all its lanes are done
        const unsigned idle = __ballot_sync(0xffffffffu, mode == kFogIdle && !needRay);
        if (idle == 0xffffffffu) {
            if (drained) break;
            unsigned item = 0;
            if (lane == 0) item = vt.tile0 + atomicAdd(queue, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= vt.tile1) { drained = true; continue; }
            if (vt.only && !vt.only[item - vt.tile0]) continue;
            uint32_t px, py;
            if (ticketToPixel(tm, item * 32u + lane, px, py)) {
                pix = size_t(py) * tm.width + px;
                k = 0; n = 2ull * p.sub * pix; needRay = true;
            }
        }
        bool fin = false;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);                   // bg.a = bg.r = bg.g = bg.b = 0 (:1020): a sample that misses the bbox
        // (1b) the next sample of the lane's pixel: the centre first, then the jittered offsets of LevelSetRayTracer (:907-909)
        if (needRay) {
            needRay = false;
            const bool first = k == 0;
            cameraRay(cam, uint32_t(pix % tm.width), uint32_t(pix / tm.width), first ? 0.5 : p.jitter[n & 15], first ? 0.5 : p.jitter[(n + 1) & 15], ray);
            if (!first) n += 2;
            if (COUNT) ++c.rays;
            worldToIndex(g, ray);
            if (clipRay(ray, g, 1)) {                                    // mPrimary->setWorldRay(pRay) (:1022)
                Tx = Ty = Tz = 1.0; Lx = Ly = Lz = 0.0;
                walk.begin(ray); mode = kFogPrimary; pendExp = 0; span = 0;
            } else fin = true;                                           // `continue` (:1022): the sample stays (0,0,0,0)
        }
#pragma unroll 1
        for (;;) {
            __syncwarp();
            // for (pT = pStep*ceil(t0/pStep); pT <= pT1; pT += pStep): past the end of a closed span -> look for the next one
            if (span == 2 && !(tcur <= tend)) span = 0;
            const bool busy = mode != kFogIdle && !pendExp;
            const bool marching = busy && (span == 2 || (span == 1 && tcur <= walk.bound && (walk.bound - walk.ts0) > 1e-9));
            const bool walking = busy && !marching;
            const int nW = __popc(__ballot_sync(0xffffffffu, walking)), nM = __popc(__ballot_sync(0xffffffffu, marching));
            const int nE = __popc(__ballot_sync(0xffffffffu, pendExp != 0));
            const bool runW = nW >= kFogBatch || (nM < kFogBatch && nE < kFogBatch);
            const bool runM = nM >= kFogBatch || (nW < kFogBatch && nE < kFogBatch);
            const bool runE = nE >= kFogBatch || (nW < kFogBatch && nM < kFogBatch);
            bool lum = false;
            // (2) walk: one unit of VolumeHDDA::hits for the active ray
            if (walking && runW) {
                double a, b;
                const int r = walk.template advance<COUNT>(g, root, sm, mode == kFogPrimary ? 0 : 2, accW, ray, a, b, c);
                if (r == kSpanEmit) { tend = b; span = 2; }                  // the open span closed at b (valid: b - a > 1e-9)
                else if (r == kSpanDone) { if (mode == kFogPrimary) fin = true; else lum = true; }   // shadow spans exhausted: Luminance
                else if (span == 1 && walk.ts0 < 0.0) span = 0;              // it closed, but too short to count (TimeSpan::valid)
                if (span == 0 && r != kSpanDone && (walk.ts0 >= 0.0 || r == kSpanEmit)) {
                    // a span opened at a: first sample time pStep*ceil(t0/pStep) (:1030-1032, :1045-1047)
                    const double st = mode == kFogPrimary ? p.pstep : p.sstep;
                    const double t0s = r == kSpanEmit ? a : walk.ts0;
                    tcur = st * ceil(t0s / st);
                    span = r == kSpanEmit ? 2 : 1;
                }
            }
            __syncwarp();
            // (3) sample: density at the current march time of the active ray
            if (marching && runM) {
                // getWorldPos(t) = indexToWorld(ray(t)); sampler.wsSample -> worldToIndex -> BoxSampler (:1034-1035,1048)
                double wx = ray.ex + ray.dx * tcur, wy = ray.ey + ray.dy * tcur, wz = ray.ez + ray.dz * tcur;
                indexToWorldPos(g, wx, wy, wz);
                const double d = boxSampleWorld<LEAF>(g, root, accV, wx, wy, wz);
                if (COUNT) { if (mode == kFogPrimary) ++c.psamples; else ++c.ssamples; }
                if (d < p.cutoff) tcur += mode == kFogPrimary ? p.pstep : p.sstep;        // continue
                else { dens = d; pendExp = mode; }
            }
            __syncwarp();
            // (4) exp: dT = Exp(extinction*density*pStep) (:1037) or sTrans *= Exp(extinction*d*sStep/(1+sT*sGain)) (:1053)
            if (pendExp && runE) {
                const bool prim = pendExp == kFogPrimary;
                pendExp = 0;
                const double den = 1.0 + tcur * p.gain;
                const double ax = prim ? p.ext[0] * dens * p.pstep : p.ext[0] * dens * p.sstep / den;
                const double ay = prim ? p.ext[1] * dens * p.pstep : p.ext[1] * dens * p.sstep / den;
                const double az = prim ? p.ext[2] * dens * p.pstep : p.ext[2] * dens * p.sstep / den;
                const double ex = exp(ax), ey = exp(ay), ez = exp(az);
                if (prim) {
                    dTx = ex; dTy = ey; dTz = ez;
                    Sx = Sy = Sz = 1.0;
                    // sRay.setEye(pPos); mShadow->setWorldRay(sRay) (:1039-1040)
                    double wx = ray.ex + ray.dx * tcur, wy = ray.ey + ray.dy * tcur, wz = ray.ez + ray.dz * tcur;
                    indexToWorldPos(g, wx, wy, wz);
                    worldToIndexPos(g, wx, wy, wz);
                    Ray sRay;
                    sRay.ex = wx; sRay.ey = wy; sRay.ez = wz;
                    sRay.dx = sm.sbase[0]; sRay.dy = sm.sbase[1]; sRay.dz = sm.sbase[2];
                    sRay.ix = sm.sbase[3]; sRay.iy = sm.sbase[4]; sRay.iz = sm.sbase[5];
                    sRay.t0 = sm.sbase[6]; sRay.t1 = sm.sbase[7];
                    if (COUNT) ++c.srays;
                    if (!clipRay(sRay, g, 1)) tcur += p.pstep;              // `continue`: no luminance for this sample
                    else {
                        // suspend the primary walk and ray, switch the lane to the shadow ray
                        sm.park(4, walk.cur);
                        sm.dt0[tid] = walk.cur.t0; sm.ts0[tid] = walk.ts0; sm.topT1[tid] = walk.topT1; sm.tcur[tid] = tcur; sm.tend[tid] = tend;
                        sm.bound[tid] = walk.bound;
                        sm.misc[tid] = (walk.lvl + 1) | (walk.needStep ? 256 : 0) | (span << 12);
                        sm.ray[0][tid] = ray.ex; sm.ray[1][tid] = ray.ey; sm.ray[2][tid] = ray.ez; sm.ray[3][tid] = ray.dx; sm.ray[4][tid] = ray.dy;
                        sm.ray[5][tid] = ray.dz;
                        ray = sRay;
                        walk.begin(ray);
                        mode = kFogShadow; span = 0;
                    }
                } else {
                    Sx *= ex; Sy *= ey; Sz *= ez;
                    if (Sx * Sx + Sy * Sy + Sz * Sz < p.cutoff) lum = true;                      // goto Luminance (:1054)
                    else tcur += p.sstep;
                }
            }
            // (5) Luminance (:1057-1060): back to the primary ray
            if (lum) {
                Lx += p.albedo[0] * Sx * Tx * (1.0 - dTx); Ly += p.albedo[1] * Sy * Ty * (1.0 - dTy); Lz += p.albedo[2] * Sz * Tz * (1.0 - dTz);
                Tx *= dTx; Ty *= dTy; Tz *= dTz;
                if (Tx * Tx + Ty * Ty + Tz * Tz < p.cutoff) fin = true;                         // goto Pixel
                else {
                    sm.unpark(4, walk.cur);
                    walk.cur.t0 = sm.dt0[tid]; walk.ts0 = sm.ts0[tid]; walk.topT1 = sm.topT1[tid]; tend = sm.tend[tid];
                    walk.bound = sm.bound[tid];
                    tcur = sm.tcur[tid] + p.pstep;
                    const int m = sm.misc[tid];
                    walk.lvl = (m & 255) - 1; walk.needStep = (m & 256) != 0; span = m >> 12;
                    ray.ex = sm.ray[0][tid]; ray.ey = sm.ray[1][tid]; ray.ez = sm.ray[2][tid]; ray.dx = sm.ray[3][tid]; ray.dy = sm.ray[4][tid];
                    ray.setDir(ray.dx, ray.dy, sm.ray[5][tid]);          // invDir = 1/dir, the same division as at ray set-up
                    mode = kFogPrimary;
                }
            }
            // (6) Pixel (:1063-1067); with more than one sample the results are summed in order and scaled by 1/samples
            if (fin) {
                if (mode != kFogIdle) {
                    out = make_float4(float(Lx), float(Ly), float(Lz), float(1.0f - (Tx + Ty + Tz) / 3.0f));
                    if (COUNT && out.w > 0.f) ++c.hits;
                }
                if (k == 0) acc = out;
                else { acc.x += out.x; acc.y += out.y; acc.z += out.z; acc.w += out.w; }
                mode = kFogIdle; pendExp = 0;
                if (++k > p.sub) film[pix] = make_float4(acc.x * p.frac, acc.y * p.frac, acc.z * p.frac, acc.w * p.frac);
                else needRay = true;
            }
                fin = false;
            out = make_float4(0.f, 0.f, 0.f, 0.f);
            if (__any_sync(0xffffffffu, needRay) || !__any_sync(0xffffffffu, mode != kFogIdle)) break;
        }
    }
    if (COUNT) flushCounters(c, counters);
}

// Film::RGBA::over (tools/RayTracer.h:252-259), one pixel per thread: 32 B read + 16 B written per pixel, HBM-bound
__global__ void k_film_over(float4* __restrict__ top, const float4* __restrict__ bottom, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float4 t = top[i], b = bottom[i];
    const float s = b.w * (1.0f - t.w);
    top[i] = make_float4(t.w * t.x + s * b.x, t.w * t.y + s * b.y, t.w * t.z + s * b.z, t.w + s);
}

// ------------------------------------------------------------------------------------------------------------
// RootNode::evalActiveBoundingBox(bbox, /*visitVoxels=*/false) (tree/RootNode.h:1532-1541, InternalNode.h:1246-1256,
// LeafNode.h:1505-1517) over the NanoVDB node arrays: one thread per (root tile, upper slot).
// out[0..2] = min, out[3..5] = max (initialised to INT_MAX / INT_MIN by the host)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void expandBox(int* out, int x, int y, int z, int dim)
{
    atomicMin(out + 0, x); atomicMin(out + 1, y); atomicMin(out + 2, z);
    atomicMax(out + 3, x + dim - 1); atomicMax(out + 4, y + dim - 1); atomicMax(out + 5, z + dim - 1);
}

// DevGrid::lowmask: one thread per mask word of every lower node
__global__ void k_build_lowmask(const uint8_t* __restrict__ base, unsigned long long lowerOff, uint32_t lowerCount, unsigned long long* __restrict__ out)
{
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const uint32_t node = uint32_t(t >> 6), w = uint32_t(t & 63u);
    if (node >= lowerCount) return;
    const uint8_t* l = base + lowerOff + (unsigned long long)node * 33856ull;
    out[t] = ldg64(l + kLowerCMask + 8u * w) | ldg64(l + kLowerVMask + 8u * w);
}

// DevGrid::halo: one 256-thread block per leaf (grid-stride), three of the 729 values per thread.  Values outside the leaf are
// what ValueAccessor::getValue returns there (neighbour leaf, tile of any level, background): a full descent from the root.
__global__ void __launch_bounds__(256) k_build_halo(const __grid_constant__ DevGrid g, unsigned long long leafOff, float* __restrict__ out)
{
    __shared__ RootSmem root;
    stageRoot(g, root);
    __syncthreads();
    for (uint32_t leaf = blockIdx.x; leaf < g.leaf_count; leaf += gridDim.x) {
        const uint8_t* lf = g.base + leafOff + (unsigned long long)leaf * 2144ull;
        const int ox = int(ldg32(lf)) & ~7, oy = int(ldg32(lf + 4)) & ~7, oz = int(ldg32(lf + 8)) & ~7;   // mBBoxMin -> origin (NanoVDB.h:4489)
        float* dst = out + size_t(leaf) * kHaloStride;
        for (uint32_t t = threadIdx.x; t < 729u; t += 256u) {
            const uint32_t lx = t / 81u, ly = (t / 9u) % 9u, lz = t % 9u;
            float v;
            if (lx < 8u && ly < 8u && lz < 8u) v = ldgf(lf + kLeafValues + 4u * ((lx << 6) | (ly << 3) | lz));
            else { TreeCursor c; c.reset(); v = c.getValue(g, root, ox + int(lx), oy + int(ly), oz + int(lz)); }
            dst[t] = v;
        }
    }
}

// The same pass validates the tree's links: a child offset must land on a node of the right level inside the buffer
// (nodes of one level are stored contiguously, NanoVDB.h:67-122) before anything is read through it; out[6] counts the bad ones.
struct NodeAreas { unsigned long long upper0, upperN, lower0, lowerN, leaf0, leafN, leafBytes; };      // byte offset of node 0 and node count, per level; size of a leaf
__device__ __forceinline__ bool inArea(const uint8_t* base, const uint8_t* p, unsigned long long first, unsigned long long count, unsigned long long size)
{
    const unsigned long long off = (unsigned long long)(p - base);
    return p >= base + first && off - first < count * size && (off - first) % size == 0ull;
}

// DevGrid::halo for quantised leaves (see the leaf kinds in vdbrt_device.cuh): per leaf 8 {minimum, quantum} pairs -- one per source
// block (own leaf, then the blocks at +z, +y, +yz, +x, +xz, +xy, +xyz) -- and the 729 CODES.  A source block that is a leaf gives its
// own pair and codes; one that is a tile or the background gives {value, 0} and code 0 (0 * 0 + value is the value, exactly).
template<int LEAF>
__global__ void __launch_bounds__(256) k_build_halo_q(const __grid_constant__ DevGrid g, unsigned long long leafOff, uint8_t* __restrict__ out)
{
    __shared__ RootSmem root;
    __shared__ const uint8_t* src[8];      // the leaf of source block r, or null
    stageRoot(g, root);
    __syncthreads();
    for (uint32_t leaf = blockIdx.x; leaf < g.leaf_count; leaf += gridDim.x) {
        const uint8_t* lf = g.base + leafOff + (unsigned long long)leaf * LeafKind<LEAF>::bytes;
        const int ox = int(ldg32(lf)) & ~7, oy = int(ldg32(lf + 4)) & ~7, oz = int(ldg32(lf + 8)) & ~7;
        uint8_t* dst = out + size_t(leaf) * LeafKind<LEAF>::block;
        if (threadIdx.x < 8) {
            const uint32_t r = threadIdx.x;
            TreeCursor c; c.reset();
            const int x = ox + int((r >> 2) << 3), y = oy + int(((r >> 1) & 1u) << 3), z = oz + int((r & 1u) << 3);
            const int depth = c.descend(g, root, x, y, z);
            float mn, q = 0.f;
            if (depth == 0) { const uint8_t* nl = TreeCursor::node(g, c.n0); src[r] = nl; mn = ldgf(nl + kLeafMinimum); q = ldgf(nl + kLeafQuantum); }
            else { src[r] = nullptr; c.template valueAt<LEAF>(g, root, depth, x, y, z, mn); }
            reinterpret_cast<float2*>(dst)[r] = make_float2(mn, q);
        }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < 729u; t += 256u) {
            const uint32_t lx = t / 81u, ly = (t / 9u) % 9u, lz = t % 9u;
            const uint8_t* nl = src[((lx >> 3) << 2) | ((ly >> 3) << 1) | (lz >> 3)];
            const uint32_t n = ((lx & 7u) << 6) | ((ly & 7u) << 3) | (lz & 7u);
            if (LEAF == kLeafFp8) dst[kQHaloCodes + t] = nl ? __ldg(nl + kLeafValues + n) : uint8_t(0);
            else reinterpret_cast<unsigned short*>(dst + kQHaloCodes)[t] = nl ? __ldg(reinterpret_cast<const unsigned short*>(nl + kLeafValues) + n) : (unsigned short)0;
        }
        __syncthreads();
    }
}

__global__ void k_node_bbox(const uint8_t* __restrict__ base, unsigned long long rootOff, uint32_t tableSize, int* out, const __grid_constant__ NodeAreas ar)
{
    const unsigned long long gid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const uint32_t tile = uint32_t(gid >> 15), n = uint32_t(gid & 32767u);
    if (tile >= tableSize) return;
    const uint8_t* t = base + rootOff + kRootTiles + kTileSize * tile;
    const unsigned long long key = ldg64(t);
    const int ox = int(uint32_t((key >> 42) & 0x1FFFFFull) << 12), oy = int(uint32_t((key >> 21) & 0x1FFFFFull) << 12), oz = int(uint32_t(key & 0x1FFFFFull) << 12);
    const long long child = ldgs64(t + 8);
    if (child == 0) { if (n == 0 && ldg32(t + 16)) expandBox(out, ox, oy, oz, 4096); return; }
    const uint8_t* u = base + rootOff + child;
    if (!inArea(base, u, ar.upper0, ar.upperN, 270400ull)) { if (n == 0) atomicAdd(out + 6, 1); return; }
    const int ux = ox + int((n >> 10) << 7), uy = oy + int(((n >> 5) & 31u) << 7), uz = oz + int((n & 31u) << 7);
    if (!maskBit(u + kUpperCMask, n)) { if (maskBit(u + kUpperVMask, n)) expandBox(out, ux, uy, uz, 128); return; }
    const uint8_t* l = u + ldgs64(u + kUpperTable + 8u * n);
    if (!inArea(base, l, ar.lower0, ar.lowerN, 33856ull)) { atomicAdd(out + 6, 1); return; }
    int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {int(0x80000000), int(0x80000000), int(0x80000000)};
    int bad = 0;
    for (uint32_t w = 0; w < 64; ++w) {
        const unsigned long long cm = ldg64(l + kLowerCMask + 8u * w), vm = ldg64(l + kLowerVMask + 8u * w);
        unsigned long long tiles = vm & ~cm, kids = cm;
        while (tiles) {
            const uint32_t m = w * 64u + uint32_t(__ffsll((long long)tiles) - 1); tiles &= tiles - 1;
            const int x = ux + int((m >> 8) << 3), y = uy + int(((m >> 4) & 15u) << 3), z = uz + int((m & 15u) << 3);
            mn[0] = min(mn[0], x); mn[1] = min(mn[1], y); mn[2] = min(mn[2], z); mx[0] = max(mx[0], x + 7); mx[1] = max(mx[1], y + 7); mx[2] = max(mx[2], z + 7);
        }
        while (kids) {
            const uint32_t m = w * 64u + uint32_t(__ffsll((long long)kids) - 1); kids &= kids - 1;
            const uint8_t* lf = l + ldgs64(l + kLowerTable + 8u * m);
            if (!inArea(base, lf, ar.leaf0, ar.leafN, ar.leafBytes)) { ++bad; continue; }
            unsigned long long any = 0;
            for (int q = 0; q < 8; ++q) any |= ldg64(lf + kLeafVMask + 8 * q);
            if (!any) continue;
            const int x = ux + int((m >> 8) << 3), y = uy + int(((m >> 4) & 15u) << 3), z = uz + int((m & 15u) << 3);
            mn[0] = min(mn[0], x); mn[1] = min(mn[1], y); mn[2] = min(mn[2], z); mx[0] = max(mx[0], x + 7); mx[1] = max(mx[1], y + 7); mx[2] = max(mx[2], z + 7);
        }
    }
    if (bad) atomicAdd(out + 6, bad);
    if (mn[0] <= mx[0]) {
        atomicMin(out + 0, mn[0]); atomicMin(out + 1, mn[1]); atomicMin(out + 2, mn[2]);
        atomicMax(out + 3, mx[0]); atomicMax(out + 4, mx[1]); atomicMax(out + 5, mx[2]);
    }
}

} // namespace vdbrt
