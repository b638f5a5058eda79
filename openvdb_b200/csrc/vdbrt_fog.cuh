// vdbrt_fog.cuh -- VolumeRender::operator() (tools/RayTracer.h:991-1070) as a WAVEFRONT of three homogeneous kernels.
//
// The reference marches a primary ray and, at every dense sample, a shadow ray towards the light, all in one loop.  On a GPU
// that loop (k_render_volume, vdbrt_kernels.cuh) makes every lane a two-mode state machine -- 153 registers, 3 CTAs per SM, 11.7
// of 32 lanes active, because the lanes of a warp are on primary or shadow rays, walking or sampling, at different times.
// But nothing on the primary ray depends on what its shadow rays find: the shadow transmittance S_k only enters the luminance
//     pLumi += albedo * S_k * pTrans_k * (1 - dT_k)        (tools/RayTracer.h:1058)
// while the march goes on with pTrans *= dT_k, and stops on pTrans alone (:1059-1060).  So:
//   1. k_fog_primary  one lane per primary ray: span walk + density samples + exp; every dense sample whose shadow ray enters the
//                     bbox (:1040) becomes a 96-byte RECORD {pTrans_k, dT_k, shadow ray} in HBM, chained per ray in sample order;
//   2. k_fog_shadow   one lane per RECORD: the shadow march (:1041-1056) -> S_k -> term_k = albedo*S_k*pTrans_k*(1-dT_k), written back;
//   3. k_fog_resolve  one thread per pixel: pLumi = ((term_0) + term_1) + ... in sample order, alpha from the final pTrans (:1063-1067),
//                     samples of a pixel summed in order, film written.
// Same samples, same operands, same order of every floating-point operation as the one-loop kernel (which stays: it counts the
// work for the roofline, and it re-renders the tiles of a batch whose records did not fit).  What the wavefront buys is SIMT
// efficiency and occupancy; what it costs is ~100 bytes of HBM traffic per dense sample -- on a path whose kernels leave HBM idle.
#pragma once
#include "vdbrt_kernels.cuh"
#ifndef VDBRT_FOG_SHADOW_BLOCKS
#define VDBRT_FOG_SHADOW_BLOCKS 6
#endif
#ifndef VDBRT_FOG_PRIMARY_BLOCKS
#define VDBRT_FOG_PRIMARY_BLOCKS 4
#endif

namespace vdbrt {

constexpr uint32_t kNoRec = 0xffffffffu;
struct FogRec {                      // one dense primary sample
    double a[3];                     // in: pTrans before the sample;  out (k_fog_shadow): the luminance term of the sample
    double dT[3];
    double eye[3], t0, t1;           // shadow ray in index space, clipped (its direction is the same for all: VolParams::light)
    uint32_t next, pad;              // next record of the same primary ray
};
static_assert(sizeof(FogRec) == 96, "FogRec layout");
struct FogRayRec { double T[3]; uint32_t head, state; };   // per primary ray: final pTrans, first record; state 0 = missed the bbox, 1 = marched
struct FogWave {
    FogRec* recs; FogRayRec* rays; uint8_t* tileFlag;      // tileFlag[t - tile0] = 1: a ray of the tile ran out of record space
    unsigned int* ctl;                                     // [0] tile queue, [1] records used, [2] flagged tiles, [3] shadow queue
    uint32_t cap;                                          // records available
    uint32_t tile0, tile1;                                 // this batch: work items [tile0, tile1) of the launch's TileMap
    uint32_t refill;                                       // shadow kernel: idle lanes that trigger a refill from the record queue (32: whole tickets)
};

// Parking area for the suspended parent levels of ONE span walk per lane (two slots: root level, upper level)
template<int THREADS>
struct FogPark {
    double t1[2][THREADS], nx[2][THREADS], ny[2][THREADS], nz[2][THREADS];
    int vx[2][THREADS], vy[2][THREADS], vz[2][THREADS];
    __device__ __forceinline__ void park(int slot, const Dda& d)
    {
        const int t = threadIdx.x;
        t1[slot][t] = d.t1; nx[slot][t] = d.nx; ny[slot][t] = d.ny; nz[slot][t] = d.nz; vx[slot][t] = d.vx; vy[slot][t] = d.vy; vz[slot][t] = d.vz;
    }
    __device__ __forceinline__ void unpark(int slot, Dda& d) const
    {
        const int t = threadIdx.x;
        d.t1 = t1[slot][t]; d.nx = nx[slot][t]; d.ny = ny[slot][t]; d.nz = nz[slot][t]; d.vx = vx[slot][t]; d.vy = vy[slot][t]; d.vz = vz[slot][t];
    }
};

// One lane's march along one ray: the lazy span walk of k_render_volume (vdbrt_kernels.cuh) -- a sample is taken as soon as the
// entry time of the last probed cell proves that it lies inside the still-open span.
struct FogMarch {
    Ray ray; SpanWalk walk;
    double tcur, tend;
    int span;                // 0: looking for a span, 1: inside an open span (end >= walk.bound), 2: span closed at tend
    __device__ __forceinline__ void begin() { walk.begin(ray); span = 0; tcur = 0.0; tend = 0.0; }
    __device__ __forceinline__ void idle()
    {
        ray.ex = ray.ey = ray.ez = 0.0; ray.setDir(1.0, 1.0, 1.0); ray.t0 = ray.t1 = 0.0;
        walk.begin(ray); walk.lvl = -1;
        walk.cur.t0 = walk.cur.t1 = walk.cur.nx = walk.cur.ny = walk.cur.nz = 0.0; walk.cur.vx = walk.cur.vy = walk.cur.vz = 0;
        tcur = tend = 0.0; span = 0;
    }
    // for (t = step*ceil(t0/step); t <= t1; t += step): past the end of a closed span -> look for the next one
    __device__ __forceinline__ void closeIfPast() { if (span == 2 && !(tcur <= tend)) span = 0; }
    __device__ __forceinline__ bool marching() const { return span == 2 || (span == 1 && tcur <= walk.bound && (walk.bound - walk.ts0) > 1e-9); }
    // one unit of VolumeHDDA::hits; returns true when the spans are exhausted
    template<class SM>
    __device__ __forceinline__ bool walkUnit(const DevGrid& g, const RootSmem& root, SM& sm, TreeCursor& acc, double step, Counters& c)
    {
        double a, b;
        const int r = walk.template advance<false>(g, root, sm, 0, acc, ray, a, b, c);
        if (r == kSpanEmit) { tend = b; span = 2; }                  // the open span closed at b (valid: b - a > 1e-9)
        else if (r == kSpanDone) return true;
        else if (span == 1 && walk.ts0 < 0.0) span = 0;              // it closed, but too short to count (TimeSpan::valid)
        if (span == 0 && (walk.ts0 >= 0.0 || r == kSpanEmit)) {
            // a span opened at a: first sample time step*ceil(t0/step) (:1030-1032, :1045-1047)
            const double t0s = r == kSpanEmit ? a : walk.ts0;
            tcur = step * ceil(t0s / step);
            span = r == kSpanEmit ? 2 : 1;
        }
        return false;
    }
};

#ifndef VDBRT_FOG_BATCH
#define VDBRT_FOG_BATCH 4
#endif
constexpr int kFogWaveBatch = VDBRT_FOG_BATCH;       // a phase runs when this many lanes want it, or when the other phases are starved

// ---------------------------------------------------------------------------------------------------------------------------
// 1. primary rays
// ---------------------------------------------------------------------------------------------------------------------------
template<int LEAF = kLeafFloat>
__global__ void __launch_bounds__(kBlockThreads, VDBRT_FOG_PRIMARY_BLOCKS)
k_fog_primary(const __grid_constant__ DevGrid g, const __grid_constant__ DevCamera cam, const __grid_constant__ VolParams p,
              const __grid_constant__ TileMap tm, const __grid_constant__ FogWave fw)
{
    __shared__ RootSmem root;
    __shared__ FogPark<kBlockThreads> sm;
    stageRoot(g, root);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const uint32_t spp = p.sub + 1u;
#ifdef VDBRT_FOG_PRIMARY_TWO_CURSORS
    TreeCursor accW, accV;               // walker cursor, sampler cursor (the reference keeps separate accessors too)
    accW.reset(); accV.reset();
#else
    TreeCursor accW; accW.reset();       // one cursor for both (see k_fog_shadow)
    TreeCursor& accV = accW;
#endif
    Counters c = {};
    bool busy = false, pendExp = false, needRay = false, drained = false;
    size_t pix = 0;
    uint32_t k = 0, rid0 = 0, prev = kNoRec, head = kNoRec, tileSlot = 0;
    unsigned long long n = 0;            // jitter index, as in LevelSetRayTracer::operator() (tools/RayTracer.h:903-913)
    FogMarch m; m.idle();
    double dens = 0.0;
    double Tx = 1.0, Ty = 1.0, Tz = 1.0;

    for (;;) {
        __syncwarp();
        // (1) a warp takes a fresh 8x4 tile when all its lanes are done
        const unsigned idle = __ballot_sync(0xffffffffu, !busy && !needRay);
        if (idle == 0xffffffffu) {
            if (drained) break;
            unsigned item = 0;
            if (lane == 0) item = fw.tile0 + atomicAdd(fw.ctl + 0, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= fw.tile1) { drained = true; continue; }
            uint32_t px, py;
            tileSlot = item - fw.tile0;
            if (ticketToPixel(tm, item * 32u + lane, px, py)) {
                pix = size_t(py) * tm.width + px;
                k = 0; n = 2ull * p.sub * pix; needRay = true;
                rid0 = (tileSlot * 32u + lane) * spp;
            }
        }
        bool fin = false, missed = false;
        // (1b) the next sample of the lane's pixel: the centre first, then the jittered offsets of LevelSetRayTracer (:907-909)
        if (needRay) {
            needRay = false;
            const bool first = k == 0;
            cameraRay(cam, uint32_t(pix % tm.width), uint32_t(pix / tm.width), first ? 0.5 : p.jitter[n & 15], first ? 0.5 : p.jitter[(n + 1) & 15], m.ray);
            if (!first) n += 2;
            worldToIndex(g, m.ray);
            Tx = Ty = Tz = 1.0; prev = kNoRec; head = kNoRec;
            if (clipRay(m.ray, g, 1)) { m.begin(); busy = true; pendExp = false; }       // mPrimary->setWorldRay(pRay) (:1022)
            else { fin = true; missed = true; }                                          // `continue` (:1022): the sample stays (0,0,0,0)
        }
#pragma unroll 1
        for (;;) {
            __syncwarp();
            m.closeIfPast();
            const bool act = busy && !pendExp;
            const bool marching = act && m.marching();
            const bool walking = act && !marching;
            bool runW = true, runM = true, runE = true;
            if (kFogWaveBatch > 1) {
                const int nW = __popc(__ballot_sync(0xffffffffu, walking)), nM = __popc(__ballot_sync(0xffffffffu, marching));
                const int nE = __popc(__ballot_sync(0xffffffffu, pendExp));
                runW = nW >= kFogWaveBatch || (nM < kFogWaveBatch && nE < kFogWaveBatch);
                runM = nM >= kFogWaveBatch || (nW < kFogWaveBatch && nE < kFogWaveBatch);
                runE = nE >= kFogWaveBatch || (nW < kFogWaveBatch && nM < kFogWaveBatch);
            }
            // (2) walk: one unit of VolumeHDDA::hits
            if (walking && runW) { if (m.walkUnit(g, root, sm, accW, p.pstep, c)) fin = true; }
            __syncwarp();
            // (3) sample: density at the current march time
            if (marching && runM) {
                // getWorldPos(t) = indexToWorld(ray(t)); sampler.wsSample -> worldToIndex -> BoxSampler (:1034-1035)
                double wx = m.ray.ex + m.ray.dx * m.tcur, wy = m.ray.ey + m.ray.dy * m.tcur, wz = m.ray.ez + m.ray.dz * m.tcur;
                indexToWorldPos(g, wx, wy, wz);
                const double d = boxSampleWorld<LEAF>(g, root, accV, wx, wy, wz);
                if (d < p.cutoff) m.tcur += p.pstep;                                      // continue (:1036)
                else { dens = d; pendExp = true; }
            }
            __syncwarp();
            // (4) exp: dT = Exp(extinction*density*pStep) (:1037); the shadow ray of the sample (:1039-1040)
            bool emit = false;
            double dTx = 1.0, dTy = 1.0, dTz = 1.0, sx = 0.0, sy = 0.0, sz = 0.0, s0 = 0.0, s1 = 0.0;
            if (pendExp && runE) {
                pendExp = false;
                dTx = exp(p.ext[0] * dens * p.pstep); dTy = exp(p.ext[1] * dens * p.pstep); dTz = exp(p.ext[2] * dens * p.pstep);
                double wx = m.ray.ex + m.ray.dx * m.tcur, wy = m.ray.ey + m.ray.dy * m.tcur, wz = m.ray.ez + m.ray.dz * m.tcur;
                indexToWorldPos(g, wx, wy, wz);
                worldToIndexPos(g, wx, wy, wz);
                Ray sRay;
                sRay.ex = wx; sRay.ey = wy; sRay.ez = wz;
                sRay.dx = p.sb[0]; sRay.dy = p.sb[1]; sRay.dz = p.sb[2];
                sRay.ix = p.sb[3]; sRay.iy = p.sb[4]; sRay.iz = p.sb[5];
                sRay.t0 = p.sb[6]; sRay.t1 = p.sb[7];
                if (!clipRay(sRay, g, 1)) m.tcur += p.pstep;                              // `continue`: no luminance, pTrans unchanged
                else { emit = true; sx = wx; sy = wy; sz = wz; s0 = sRay.t0; s1 = sRay.t1; }
            }
            __syncwarp();
            // (5) the sample's record (one atomic per warp), then what the reference does after Luminance: pTrans *= dT, cut-off test
            const unsigned em = __ballot_sync(0xffffffffu, emit);
            if (em) {
                unsigned base = 0;
                if (lane == __ffs(em) - 1) base = atomicAdd(fw.ctl + 1, (unsigned)__popc(em));
                base = __shfl_sync(0xffffffffu, base, __ffs(em) - 1);
                if (emit) {
                    const unsigned idx = base + __popc(em & ((1u << lane) - 1u));
                    if (idx >= fw.cap) {
                        // no room: the one-loop kernel renders this tile again (k_render_volume with the flags)
                        if (!fw.tileFlag[tileSlot]) { fw.tileFlag[tileSlot] = 1; atomicAdd(fw.ctl + 2, 1u); }
                        busy = false; needRay = false; fin = false; k = spp;
                    } else {
                        FogRec r;
                        r.a[0] = Tx; r.a[1] = Ty; r.a[2] = Tz; r.dT[0] = dTx; r.dT[1] = dTy; r.dT[2] = dTz;
                        r.eye[0] = sx; r.eye[1] = sy; r.eye[2] = sz; r.t0 = s0; r.t1 = s1; r.next = kNoRec; r.pad = 0;
                        fw.recs[idx] = r;
                        if (prev != kNoRec) fw.recs[prev].next = idx; else head = idx;
                        prev = idx;
                        Tx *= dTx; Ty *= dTy; Tz *= dTz;                                 // (:1059)
                        if (Tx * Tx + Ty * Ty + Tz * Tz < p.cutoff) fin = true;          // goto Pixel (:1060)
                        else m.tcur += p.pstep;
                    }
                }
            }
            // (6) the ray is done: its row of the ray table; next sample of the pixel
            if (fin) {
                FogRayRec rr;
                rr.T[0] = Tx; rr.T[1] = Ty; rr.T[2] = Tz; rr.head = head; rr.state = missed ? 0u : 1u;
                fw.rays[rid0 + k] = rr;
                busy = false; pendExp = false;
                if (++k <= p.sub) needRay = true;
            }
            fin = false; missed = false;
            if (__any_sync(0xffffffffu, needRay) || !__any_sync(0xffffffffu, busy)) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// 2. shadow rays: one lane per record, 32 consecutive records per warp ticket
// ---------------------------------------------------------------------------------------------------------------------------
template<int LEAF = kLeafFloat>
__global__ void __launch_bounds__(kBlockThreads, VDBRT_FOG_SHADOW_BLOCKS)
k_fog_shadow(const __grid_constant__ DevGrid g, const __grid_constant__ VolParams p, const __grid_constant__ FogWave fw)
{
    __shared__ RootSmem root;
    __shared__ FogPark<kBlockThreads> sm;
    stageRoot(g, root);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const uint32_t nRec = min(fw.ctl[1], fw.cap);
    // ONE tree cursor for the span walk and the sampler (the reference keeps two accessors): a cursor is only a cache, and the samples
    // lie in the span the walk has just crossed, mostly in the same lower node -- six registers less buy a sixth CTA per SM
    TreeCursor accW; accW.reset();
    TreeCursor& accV = accW;
    Counters c = {};
    // every ray of this kernel has the direction p.sb (kernel parameters): the six direction registers of a lane are never anything else
    FogMarch m; m.idle();
    m.ray.dx = p.sb[0]; m.ray.dy = p.sb[1]; m.ray.dz = p.sb[2]; m.ray.ix = p.sb[3]; m.ray.iy = p.sb[4]; m.ray.iz = p.sb[5];
    bool busy = false, pendExp = false, pendLum = false, drained = false;
    uint32_t rec = 0;
    double dens = 0.0, Sx = 1.0, Sy = 1.0, Sz = 1.0;

    // Lanes are re-fed one by one: a shadow ray is set up by loading its record (no camera arithmetic), and the records a warp
    // draws next lie right behind the ones it is working on -- they were appended by the same primary warps -- so, unlike in the
    // level-set kernel, feeding single lanes costs neither set-up time nor locality (fw.refill idle lanes trigger a refill).
    for (;;) {
        __syncwarp();
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (idle == 0xffffffffu && drained) break;
        if (!drained && (idle == 0xffffffffu || (unsigned)__popc(idle) >= fw.refill)) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(fw.ctl + 3, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned)__popc(idle) >= nRec) drained = true;
            const unsigned mine = base + __popc(idle & ((1u << lane) - 1u));
            if (!busy && mine < nRec) {
                rec = mine;
                const FogRec* r = fw.recs + rec;
                m.ray.ex = r->eye[0]; m.ray.ey = r->eye[1]; m.ray.ez = r->eye[2];
                m.ray.t0 = r->t0; m.ray.t1 = r->t1;
                m.begin();
                Sx = Sy = Sz = 1.0; busy = true; pendExp = false;
            }
        }
#pragma unroll 1
        for (;;) {
            __syncwarp();
            m.closeIfPast();
            const bool act = busy && !pendExp;
            const bool marching = act && m.marching();
            const bool walking = act && !marching;
            bool runW = true, runM = true, runE = true;
            if (kFogWaveBatch > 1) {
                const int nW = __popc(__ballot_sync(0xffffffffu, walking)), nM = __popc(__ballot_sync(0xffffffffu, marching));
                const int nE = __popc(__ballot_sync(0xffffffffu, pendExp));
                runW = nW >= kFogWaveBatch || (nM < kFogWaveBatch && nE < kFogWaveBatch);
                runM = nM >= kFogWaveBatch || (nW < kFogWaveBatch && nE < kFogWaveBatch);
                runE = nE >= kFogWaveBatch || (nW < kFogWaveBatch && nM < kFogWaveBatch);
            }
            bool lum = false;
            if (walking && runW) { if (m.walkUnit(g, root, sm, accW, p.sstep, c)) lum = true; }     // shadow spans exhausted: Luminance
            __syncwarp();
            if (marching && runM) {
                double wx = m.ray.ex + m.ray.dx * m.tcur, wy = m.ray.ey + m.ray.dy * m.tcur, wz = m.ray.ez + m.ray.dz * m.tcur;
                indexToWorldPos(g, wx, wy, wz);
                const double d = boxSampleWorld<LEAF>(g, root, accV, wx, wy, wz);         // (:1048)
                if (d < p.cutoff) m.tcur += p.sstep;                                      // continue (:1049)
                else { dens = d; pendExp = true; }
            }
            __syncwarp();
            // sTrans *= Exp(extinction*d*sStep/(1+sT*sGain)) (:1053)
            if (pendExp && runE) {
                pendExp = false;
                const double den = 1.0 + m.tcur * p.gain;
                Sx *= exp(p.ext[0] * dens * p.sstep / den); Sy *= exp(p.ext[1] * dens * p.sstep / den); Sz *= exp(p.ext[2] * dens * p.sstep / den);
                if (Sx * Sx + Sy * Sy + Sz * Sz < p.cutoff) lum = true;                   // goto Luminance (:1054)
                else m.tcur += p.sstep;
            }
            // the shadow ray is done: its lane stops; the sample's term is written after the loop, with those of the other finished lanes
            if (lum) { busy = false; pendExp = false; pendLum = true; }
            const unsigned running = __ballot_sync(0xffffffffu, busy);
            if (running == 0u || (!drained && 32u - (unsigned)__popc(running) >= fw.refill)) break;
        }
        // Luminance (:1058): the term of this sample, in the reference's order of multiplications
        if (pendLum) {
            pendLum = false;
            FogRec* r = fw.recs + rec;
            const double ax = p.albedo[0] * Sx * r->a[0] * (1.0 - r->dT[0]);
            const double ay = p.albedo[1] * Sy * r->a[1] * (1.0 - r->dT[1]);
            const double az = p.albedo[2] * Sz * r->a[2] * (1.0 - r->dT[2]);
            r->a[0] = ax; r->a[1] = ay; r->a[2] = az;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// 3. pixels (:1063-1067); with more than one sample the results are summed in order and scaled by 1/samples
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fog_resolve(const __grid_constant__ VolParams p, const __grid_constant__ TileMap tm, const __grid_constant__ FogWave fw, float4* __restrict__ film)
{
    const uint32_t spp = p.sub + 1u;
    const unsigned long long slots = (unsigned long long)(fw.tile1 - fw.tile0) * 32ull;
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < slots; s += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t tileSlot = uint32_t(s >> 5);
        if (fw.tileFlag[tileSlot]) continue;                     // re-rendered by the one-loop kernel
        uint32_t px, py;
        if (!ticketToPixel(tm, (fw.tile0 + tileSlot) * 32u + uint32_t(s & 31u), px, py)) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t k = 0; k < spp; ++k) {
            const FogRayRec rr = fw.rays[s * spp + k];
            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);        // bg.a = bg.r = bg.g = bg.b = 0 (:1020): a sample that misses the bbox
            if (rr.state) {
                double Lx = 0.0, Ly = 0.0, Lz = 0.0;
                for (uint32_t i = rr.head; i != kNoRec;) {
                    const FogRec* r = fw.recs + i;
                    Lx += r->a[0]; Ly += r->a[1]; Lz += r->a[2];
                    i = r->next;
                }
                out = make_float4(float(Lx), float(Ly), float(Lz), float(1.0f - (rr.T[0] + rr.T[1] + rr.T[2]) / 3.0f));
            }
            if (k == 0) acc = out;
            else { acc.x += out.x; acc.y += out.y; acc.z += out.z; acc.w += out.w; }
        }
        film[size_t(py) * tm.width + px] = make_float4(acc.x * p.frac, acc.y * p.frac, acc.z * p.frac, acc.w * p.frac);
    }
}

} // namespace vdbrt
