// icp_persist.cuh -- persistent, shared-memory-resident ICP: one thread-block cluster per (source, target) pair runs
// EVERY iteration of align() -- and, for the pairwise pipeline, the coarse AND the fine align (icp:95 + icp:111) --
// inside a single launch.  (Included by icp.cu inside its anonymous namespace.)
//
// Why: edge clouds are ~10^4 points, so a per-iteration launch is bound by dependent L2 round trips and launch
// latency, not by HBM (profiles/r01_v1_summary.md).  B200 gives 227 KB of shared memory per CTA: the whole
// voxel-filtered target (<= 14336 points as SoA x/y/z = 168 KB; the 16-bit original indices stay in global memory) and
// its cell table (4096 x 8 B) fit in one SM, so every neighbour-cell probe and candidate read is an LDS (~30 cycles)
// instead of an L2 access (~250 cycles), the convergence test never leaves the SM and the host never polls.
//   * cluster of CL CTAs per pair (CL = 4, 2 or 1, chosen PER PAIR by the host so that the launches of one batch fill
//     the chip and the largest pairs get the most CTAs): every CTA holds a full replica of the target grid and owns
//     1/CL of the source points; the 17 partial sums are PUSHED into every sibling's shared memory (DSMEM stores), one
//     cluster barrier later each CTA sums the CL partials in rank order and redundantly runs the same solve, so no
//     broadcast is needed and all CTAs of a pair take the same convergence decision.
//   * v4: the CTA's slice of the working source cloud lives in REGISTERS (P_REGP points per thread: x, y, z, certified
//     bound, cached slot), loaded once per stage; an iteration touches global memory only for the rare exact-distance
//     tie.  Slices above P_REGP * 512 points are streamed through the same registers in chunks (all loads of a chunk in
//     flight at once) and written back every iteration, exactly like PCL's input_transformed.
//   * points whose certified bound no longer decides are compacted IN POINT ORDER into a shared-memory work list
//     (block scan of per-thread counts, no atomics): item j is always re-queried by the same lanes, every fp64 sum is
//     taken in a fixed order, results are reproducible run to run.
// Exactness is unchanged: cells are 4.1 x the gate, the gate ball touches <= 2x2x2 cells, distances use the FLANN
// L2_Simple order, ties go to the lowest original index.
#pragma once
// (icp.cu includes <cooperative_groups.h> and defines `cg` before entering its anonymous namespace)

constexpr int P_THREADS = 512;
constexpr int P_NTMAX = 14336;                 // target points resident per CTA
constexpr int P_CAP = 4096;                    // cell-table slots (power of two)
constexpr unsigned P_EMPTY = 0xFFFFFFFFu;
constexpr int P_WARPS = P_THREADS / 32;
constexpr int P_G = 4;                         // lanes that share one re-query (phase B)
#ifndef RSPCL_P_REGP
#define RSPCL_P_REGP 9
#endif
constexpr int P_REGP = RSPCL_P_REGP;           // source points held in registers per thread
constexpr int P_CHUNK = P_REGP * P_THREADS;    // points per register-resident chunk of a CTA's slice
constexpr int P_WLCAP = 896;                   // work-list entries (16 B each) per phase-A/B round
constexpr int P_CLMAX = 8;                     // largest cluster (portable limit)

// MUFU.SQRT (2^-22 relative error): only used for bounds that carry a 1e-5 safety margin
__device__ __forceinline__ float sqrt_approx(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// 17-sum accumulation (global-memory path: everything in fp64)
__device__ __forceinline__ void icp_accumulate(double* acc, float px, float py, float pz, float qx, float qy, float qz,
                                               float d2) {
  const double sx = px, sy = py, sz = pz, tx = qx, ty = qy, tz = qz;
  acc[0] += 1.0;
  acc[1] += sx; acc[2] += sy; acc[3] += sz;
  acc[4] += tx; acc[5] += ty; acc[6] += tz;
  acc[7] += sx * tx; acc[8] += sx * ty; acc[9] += sx * tz;
  acc[10] += sy * tx; acc[11] += sy * ty; acc[12] += sy * tz;
  acc[13] += sz * tx; acc[14] += sz * ty; acc[15] += sz * tz;
  acc[16] += (double)d2;
}

// the same sums with the correspondence count kept as an integer: 16 doubles {s, t, s t^T, d2} + n
__device__ __forceinline__ void icp_accumulate16(double* a, int& n, float px, float py, float pz, float qx, float qy,
                                                 float qz, float d2) {
  const double sx = px, sy = py, sz = pz, tx = qx, ty = qy, tz = qz;
  n += 1;
  a[0] += sx; a[1] += sy; a[2] += sz;
  a[3] += tx; a[4] += ty; a[5] += tz;
  a[6] += sx * tx; a[7] += sx * ty; a[8] += sx * tz;
  a[9] += sy * tx; a[10] += sy * ty; a[11] += sy * tz;
  a[12] += sz * tx; a[13] += sz * ty; a[14] += sz * tz;
  a[15] += (double)d2;
}

// Warp reduction of 16 doubles with a halving butterfly: at every step a lane keeps one half of its values and sends the
// other half to its partner, so the warp issues 8 + 4 + 2 + 1 + 1 = 16 64-bit exchanges instead of 16 x 5 = 80.  On
// return lane l holds the warp total of value (l >> 1) (both lanes of a pair hold the same number).  Fixed tree: the
// result does not depend on scheduling.
__device__ __forceinline__ double warp_reduce16(const double* a, int lane) {
  double b8[8], b4[4], b2[2];
  bool up = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double send = up ? a[i] : a[i + 8], keep = up ? a[i + 8] : a[i];
    b8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = up ? b8[i] : b8[i + 4], keep = up ? b8[i + 4] : b8[i];
    b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = up ? b4[i] : b4[i + 2], keep = up ? b4[i + 2] : b4[i];
    b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = (lane & 2) != 0;
  const double send = up ? b2[0] : b2[1], keep = up ? b2[1] : b2[0];
  double r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

struct PersistSmem {
  float tx[P_NTMAX], ty[P_NTMAX], tz[P_NTMAX];
  uint2 tab[P_CAP];                 // {key, start << 16 | count}
  union {
    unsigned short fill[P_CAP];     // build-time cursors
    float4 wl[P_WLCAP];             // work list of one round: in {x, y, z, slot+1 << 16 | index in chunk}, out {new bound, -, -, new slot+1 << 16}
  };
  double red[2][P_WARPS][16];       // per-warp totals of the 16 fp64 sums: [0] phase A (cached points), [1] phase B (re-queried)
  int redn[2][P_WARPS];             // per-warp correspondence counts, same split
  int wtot[2][P_WARPS];             // per-warp work-list counts (double-buffered: one barrier per use)
  double xch[2][P_CLMAX][NRED];     // partial sums of every CTA of the cluster (st.async through DSMEM), by iteration parity
  unsigned long long mbar[2];       // transaction barriers the siblings' stores complete, by iteration parity
  double tot[NRED];
  IcpState st;                      // replicated per CTA
  float M[16];
  float T1[16];                     // final transform of the coarse stage (the fine stage re-seeds its source from it)
  int origin[3];                    // minimum cell coordinates of the target
  int cmax[3];
  int flags[4];                     // [0] build failed -> fallback, [1] dummy active counter
};
static_assert(sizeof(PersistSmem) <= 232448, "PersistSmem exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ unsigned p_hash(unsigned key) { return (key * 0x9E3779B1u) >> (32 - 12); }  // log2(P_CAP) = 12
static_assert(P_CAP == 4096, "p_hash assumes 4096 slots");

__device__ __forceinline__ int p_cell(float v, float inv_cs) { return __float2int_rd(__fmul_rn(v, inv_cs)); }

// Arguments that do not fit comfortably in a parameter list
struct PersistArgs {
  const float4* src;        // original source batch (never written)
  const int* count;         // source counts
  int sstride;
  float4* work;             // streamed slices only: working cloud (.w = cached slot + 1) and bounds between iterations
  float* lb;
  int wstride;
  IcpState* st1;            // in: state initialised by k_icp_init (guess); out: result of the (coarse) align
  IcpState* st2;            // out: result of the second (fine) align, run from identity on the source moved by the first
  int n_stages;             // 1 or 2
  const double* prev2;      // second align: correspondences_prev_mse_ of its PCL object on entry (null: DBL_MAX, a fresh object)
  const float4* tgt;
  const int* tcount;
  int tstride;
  int shared_target;
  float inv_cs;
  int* corr_out;            // optional correspondence dump [iteration][pair][wstride] (match index or -1), first align only
  int corr_iters;           // iterations to dump (1 = PCL-style first correspondences)
  int n_pairs_total;        // pairs of the whole batch (dump layout)
  int* status;              // 1: the pair does not fit the shared-memory grid -> global-memory path
  int* tstart;              // optional [pair]: %globaltimer >> 10 (~us) at which the pair's cluster started (wave feedback)
  const int* order;         // pairs of this launch, largest first
  unsigned short* tidx;     // scratch: original target index of every cell-sorted point, one replica per CTA of a cluster
  int tidx_rep;             // replicas per pair in tidx (>= the largest cluster of the batch)
  long long* dbg;           // optional per-CTA cycle counters
  long long* dbg_iter;      // optional per-iteration trace of the launch's first CTA: {re-queried points, phase-B cycles, iteration cycles}
};


// One pass of exact re-queries over the work list: G lanes per item.
//   G = 4: every lane probes two of the (<= 8) neighbour cells, then the four lanes walk all occupied cells together,
//          lane c taking every 4th candidate (balanced however unevenly the points are spread over the cells);
//   G >= 8 (short lists): lane c owns cell c & 7 and walks every (G/8)-th candidate of it -- the dependent chain of a
//          re-query shrinks with G, which is what bounds an iteration with only a handful of undecided points.
// The group's top-2 {winner, runner-up distance} is merged with xor shuffles; lane 0 of the group writes the result.
template <int G>
__device__ __forceinline__ void persist_requery_pass(PersistSmem& S, const unsigned short* __restrict__ TI, int nw, int wid,
                                                     int lane, float inv_cs, float rmax, float slack, int ox, int oy, int oz,
                                                     int span, int spany, int spanz, double max_dist_sqr, int* CO, int jbase,
                                                     int CL, int crank, double* acc, int& ncorr) {
  constexpr int IPW = 32 / G;  // items per warp and pass
  for (int jb = wid * IPW; jb < nw; jb += P_WARPS * IPW) {  // warp-uniform trip count (full-mask shuffles)
    const int j = jb + lane / G, c = lane % G;
    const bool active = j < nw;
    const float4 p = S.wl[active ? j : 0];
    const unsigned wbits = __float_as_uint(p.w);
    const int iloc = (int)(wbits & 0xFFFFu);
    int kp = active ? (int)(wbits >> 16) - 1 : -1;
    float bdc = INFINITY, rr = rmax;  // cached incumbent
    if (kp >= 0) {
      bdc = dist2_l2simple(p.x, p.y, p.z, S.tx[kp], S.ty[kp], S.tz[kp]);
      const float s1 = __fmaf_rn(sqrt_approx(bdc), 1.0001f, 1e-7f);
      if (s1 > rmax) {  // the cached point left the largest ball we can certify: start over
        kp = -1;
        bdc = INFINITY;
      } else {
        rr = fminf(rmax, s1 + slack);
      }
    }
    int kbest = -1;
    float bd = INFINITY, d2nd = INFINITY;  // d2nd: smallest squared distance among scanned points other than the winner
    const int x0 = p_cell(p.x - rr, inv_cs) - ox, x1 = p_cell(p.x + rr, inv_cs) - ox;
    const int y0 = p_cell(p.y - rr, inv_cs) - oy, y1 = p_cell(p.y + rr, inv_cs) - oy;
    const int z0 = p_cell(p.z - rr, inv_cs) - oz, z1 = p_cell(p.z + rr, inv_cs) - oz;
    auto probe = [&](int cc) -> unsigned {  // start << 16 | count of neighbour cell cc (0: empty / duplicate / outside)
      const bool dup = ((cc & 1) && x1 == x0) || ((cc & 2) && y1 == y0) || ((cc & 4) && z1 == z0);
      const int ix = (cc & 1) ? x1 : x0, iy = (cc & 2) ? y1 : y0, iz = (cc & 4) ? z1 : z0;
      const bool inside = (unsigned)ix <= (unsigned)span && (unsigned)iy <= (unsigned)spany && (unsigned)iz <= (unsigned)spanz;
      if (!active || dup || !inside) return 0u;
      const unsigned key = ((unsigned)ix << 20) | ((unsigned)iy << 10) | (unsigned)iz;
      unsigned s = p_hash(key);
      uint2 e = S.tab[s];
      while (e.x != key && e.x != P_EMPTY) {
        s = (s + 1) & (P_CAP - 1);
        e = S.tab[s];
      }
      return e.x == key ? e.y : 0u;
    };
    auto walk = [&](int b, int en, int first, int step) {
      for (int k = b + first; k < en; k += step) {
        const float d = (k == kp) ? INFINITY : dist2_l2simple(p.x, p.y, p.z, S.tx[k], S.ty[k], S.tz[k]);
        if (d == bd && d < INFINITY) {  // exact tie (rare): the lowest original index wins
          if (TI[k] < TI[kbest]) kbest = k;
          d2nd = bd;
          continue;
        }
        const bool lt = d < bd;
        d2nd = lt ? bd : fminf(d2nd, d);
        kbest = lt ? k : kbest;
        bd = lt ? d : bd;
      }
    };
    if (G == 4) {
      const unsigned cell0 = probe(c), cell1 = probe(c + 4);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll 1
        for (int sl = 0; sl < 4; ++sl) {
          const unsigned ey = __shfl_sync(0xffffffffu, q ? cell1 : cell0, sl, 4);
          const int b = (int)(ey >> 16);
          walk(b, b + (int)(ey & 0xFFFFu), c, 4);
        }
      }
    } else {
      const unsigned ey = probe(c & 7);
      const int b = (int)(ey >> 16);
      walk(b, b + (int)(ey & 0xFFFFu), c >> 3, G / 8);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {  // top-2 merge across the group (every lane ends with the group result)
      const float obd = __shfl_xor_sync(0xffffffffu, bd, o), o2 = __shfl_xor_sync(0xffffffffu, d2nd, o);
      const int ok_ = __shfl_xor_sync(0xffffffffu, kbest, o);
      if (obd < bd || (obd == bd && ok_ >= 0 && kbest >= 0 && TI[ok_] < TI[kbest])) {  // (the lanes' candidate sets are disjoint)
        d2nd = fminf(o2, bd);
        bd = obd;
        kbest = ok_;
      } else {
        d2nd = fminf(d2nd, obd);
      }
    }
    if (kp >= 0) {  // merge the cached incumbent (skipped by the scan)
      if (bdc < bd || (bdc == bd && kbest >= 0 && TI[kp] < TI[kbest])) {
        d2nd = bd;
        bd = bdc;
        kbest = kp;
      } else {
        d2nd = fminf(d2nd, bdc);
      }
    }
    if (active && c == 0) {
      // points outside the scanned cells are farther than rr minus the rounding of the cell-boundary test
      const float edge = rr - 2e-7f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + rr);
      float lbn = fminf(sqrt_approx(d2nd), edge) * 0.9999f;
      int cache = kbest;
      if (kbest >= 0 && !(__fmaf_rn(sqrt_approx(bd), 1.0001f, 1e-7f) < lbn)) {
        // a winner that the new bound can never confirm (beyond the certified ball, or as near as the runner-up) is not
        // worth caching: fold it into the bound over ALL target points, so that a far point is left alone until it has
        // moved instead of being re-queried every iteration
        lbn = fminf(lbn, sqrt_approx(bd) * 0.9999f);
        cache = -1;
      }
      const bool ok = kbest >= 0 && !((double)bd > max_dist_sqr);
      if (CO) CO[(jbase + iloc) * CL + crank] = ok ? (int)TI[kbest] : -1;
      if (ok) icp_accumulate16(acc, ncorr, p.x, p.y, p.z, S.tx[kbest], S.ty[kbest], S.tz[kbest], bd);
      S.wl[j] = make_float4(lbn, 0.f, 0.f, __uint_as_float((unsigned)(cache + 1)));
    }
  }
}

template <bool DBG>
__global__ void __launch_bounds__(P_THREADS, 1) k_icp_persist(const PersistArgs A, const IcpDevParams prm) {
  extern __shared__ __align__(16) unsigned char p_smem_raw[];
  PersistSmem& S = *reinterpret_cast<PersistSmem*>(p_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int CL = (int)cg::this_cluster().num_blocks();  // cluster size of this launch (1..8), chosen per pair by the host
  const int pair = A.order[blockIdx.x / CL];  // largest pairs first: later waves fill the gaps (LPT)
  const int crank = (CL > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int tseg = A.shared_target ? 0 : pair;
  const int nt = A.tcount[tseg];
  const int ns = A.count[pair];
  const float inv_cs = A.inv_cs;
  const float4* T = A.tgt + (size_t)tseg * A.tstride;
  // original target index of every cell-sorted point: only needed to break exact distance ties and to report
  // correspondences, so it lives in global memory (one copy per CTA: the order inside a cell depends on each CTA's
  // atomics) and the 2 B per point it would cost in shared memory buy 2048 more resident target points
  unsigned short* TI = A.tidx + ((size_t)pair * A.tidx_rep + crank) * P_NTMAX;
  if (A.tstart && tid == 0 && crank == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    A.tstart[pair] = (int)((gt >> 10) & 0x7fffffffull);
  }

  // ---------------------------------------------------------------- build the target replica in shared memory
  if (tid == 0) {
    S.origin[0] = S.origin[1] = S.origin[2] = INT_MAX;
    S.cmax[0] = S.cmax[1] = S.cmax[2] = INT_MIN;
    S.flags[0] = (nt > P_NTMAX) ? 1 : 0;
    S.flags[1] = 1;
    S.st = A.st1[pair];
    mbar_init(smem_u32(&S.mbar[0]), 1);
    mbar_init(smem_u32(&S.mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < P_CAP; k += P_THREADS) {
    S.tab[k] = make_uint2(P_EMPTY, 0u);
    S.fill[k] = 0;
  }
  __syncthreads();
  const bool too_big = S.flags[0] != 0;
  if (!too_big) {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = tid; i < nt; i += P_THREADS) {
      const float4 p = T[i];
      if (!finite3(p.x, p.y, p.z)) continue;
      const int c[3] = {p_cell(p.x, inv_cs), p_cell(p.y, inv_cs), p_cell(p.z, inv_cs)};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = min(mn[a], c[a]);
        mx[a] = max(mx[a], c[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
      mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
      if (lane == 0) {
        atomicMin(&S.origin[a], mn[a]);
        atomicMax(&S.cmax[a], mx[a]);
      }
    }
  }
  __syncthreads();
  if (!too_big && tid == 0) {
    for (int a = 0; a < 3; ++a)
      if (S.cmax[a] >= S.origin[a] && (long long)S.cmax[a] - S.origin[a] > 1022) S.flags[0] = 1;  // 10-bit local coordinates
  }
  __syncthreads();
  const int ox = S.origin[0], oy = S.origin[1], oz = S.origin[2];
  // pass A1: insert the occupied cells with ORDERED linear probing (the smaller key keeps the slot, the larger one is
  // carried on): the final table is the same whatever order the atomics land in
  if (!S.flags[0]) {
    for (int i = tid; i < nt; i += P_THREADS) {
      const float4 p = T[i];
      if (!finite3(p.x, p.y, p.z)) continue;
      unsigned key = ((unsigned)(p_cell(p.x, inv_cs) - ox) << 20) | ((unsigned)(p_cell(p.y, inv_cs) - oy) << 10) |
                     (unsigned)(p_cell(p.z, inv_cs) - oz);
      unsigned s = p_hash(key);
      int probes = 0;
      while (true) {
        const unsigned prev = atomicMin(&S.tab[s].x, key);  // P_EMPTY is the largest value
        if (prev == P_EMPTY || prev == key) break;
        if (prev > key) key = prev;  // displaced a larger key: it moves on
        s = (s + 1) & (P_CAP - 1);
        if (++probes >= P_CAP) {  // table full: this pair goes to the global-memory path
          S.flags[0] = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
  // pass A2: count the points of every cell
  if (!S.flags[0]) {
    for (int i = tid; i < nt; i += P_THREADS) {
      const float4 p = T[i];
      if (!finite3(p.x, p.y, p.z)) continue;
      const unsigned key = ((unsigned)(p_cell(p.x, inv_cs) - ox) << 20) | ((unsigned)(p_cell(p.y, inv_cs) - oy) << 10) |
                           (unsigned)(p_cell(p.z, inv_cs) - oz);
      unsigned s = p_hash(key);
      while (S.tab[s].x != key) s = (s + 1) & (P_CAP - 1);
      atomicAdd(&S.tab[s].y, 1u);
    }
  }
  __syncthreads();
  {  // keep the load factor <= 0.85 so that probes of absent cells always terminate quickly
    int occ = 0;
    for (int k = tid; k < P_CAP; k += P_THREADS) occ += (S.tab[k].x != P_EMPTY) ? 1 : 0;
    occ = __reduce_add_sync(0xffffffffu, occ);
    if (lane == 0) S.redn[0][wid] = occ;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < P_WARPS; ++w) t += S.redn[0][w];
      if (t > (P_CAP * 85) / 100) S.flags[0] = 1;
    }
    __syncthreads();
  }
  if (S.flags[0]) {  // uniform across the CTA (and across the cluster: every CTA builds the same replica)
    if (tid == 0 && crank == 0) A.status[pair] = 1;
    return;
  }
  // exclusive scan of the per-slot counts -> start offsets (P_CAP / P_THREADS slots per thread)
  {
    constexpr int PER = P_CAP / P_THREADS;
    int loc[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      loc[k] = (int)S.tab[tid * PER + k].y;
      sum += loc[k];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) S.redn[1][wid] = incl;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < P_WARPS; ++w) woff += (w < wid) ? S.redn[1][w] : 0;
    int excl = woff + incl - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      S.tab[tid * PER + k].y = ((unsigned)excl << 16) | (unsigned)loc[k];
      excl += loc[k];
    }
  }
  __syncthreads();
  // pass B: scatter the points into cell order.  (The order INSIDE a cell depends on how the atomics land; nothing
  // observable depends on it: distance ties are resolved through TI, and slot numbers never leave this launch.)
  for (int i = tid; i < nt; i += P_THREADS) {
    const float4 p = T[i];
    if (!finite3(p.x, p.y, p.z)) continue;
    const unsigned key = ((unsigned)(p_cell(p.x, inv_cs) - ox) << 20) | ((unsigned)(p_cell(p.y, inv_cs) - oy) << 10) |
                         (unsigned)(p_cell(p.z, inv_cs) - oz);
    unsigned s = p_hash(key);
    while (S.tab[s].x != key) s = (s + 1) & (P_CAP - 1);
    // 16-bit cursors: atomicAdd on the containing 32-bit word
    unsigned* wordp = reinterpret_cast<unsigned*>(S.fill) + (s >> 1);
    const unsigned sh = (s & 1u) * 16u;
    const unsigned old = atomicAdd(wordp, 1u << sh);
    const int pos = (int)(S.tab[s].y >> 16) + (int)((old >> sh) & 0xFFFFu);
    S.tx[pos] = p.x;
    S.ty[pos] = p.y;
    S.tz[pos] = p.z;
    TI[pos] = (unsigned short)i;
  }
  __threadfence_block();
  __syncthreads();
  // every CTA of the cluster is running (and has initialised its transaction barriers) before anyone stores into its
  // shared memory
  if (CL > 1) cg::this_cluster().sync();

  // ---------------------------------------------------------------- iterations
  const float r = prm.search_r;               // gate * 1.01
  const float rmax = __fdividef(0.485f, inv_cs);  // largest ball that touches <= 2 cells per axis (~1.99 x gate)
  const float slack = 0.5f * r;
  const int span = S.cmax[0] - ox, spany = S.cmax[1] - oy, spanz = S.cmax[2] - oz;
  // This CTA owns the source points crank, crank + CL, crank + 2 CL, ... (interleaved, so that the CTAs of a cluster see
  // the same mix of matched / unmatched / re-queried points and reach the exchange at about the same time); local index j
  // <-> source index j * CL + crank.
  const int nloc = ns / CL + (crank < ns % CL ? 1 : 0);
  const int slice_off = crank * (ns / CL) + min(crank, ns % CL);  // exclusive prefix of nloc over the ranks
  const int nchunks = (nloc + P_CHUNK - 1) / P_CHUNK;
  const bool resident = nchunks <= 1;     // the whole slice stays in registers between iterations
  const float4* SRC = A.src + (size_t)pair * A.sstride;
  float4* W = A.work + (size_t)pair * A.wstride + slice_off;  // streamed slices: CTA-local order
  float* LB = A.lb + (size_t)pair * A.wstride + slice_off;

  // the resident chunk: position, certified bound, cached slot + 1 (0: none) of local point  jbase + k * P_THREADS + tid
  float px[P_REGP], py[P_REGP], pz[P_REGP], lbv[P_REGP];
  unsigned slot[P_REGP];
#pragma unroll
  for (int k = 0; k < P_REGP; ++k) {
    px[k] = py[k] = pz[k] = NAN;
    lbv[k] = 0.f;
    slot[k] = 0u;
  }

  int parity = 0, wpar = 0, dbg_it = 0;
  unsigned mphase = 0;  // bit b: phase parity the next wait on mbar[b] expects
  // debug phase timers (DBG instantiation only: thread 0 of every CTA, RSPCL_PERSIST_DBG=1)
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tc = DBG ? clock64() : 0, n_rescan = 0;
#define P_TICK(k) do { if (DBG && tid == 0) { const long long n_ = clock64(); tph[k] += n_ - tc; tc = n_; } } while (0)
  const long long t_begin = tc;

  for (int stage = 0; stage < A.n_stages; ++stage) {
    // how a chunk enters the registers: 1 = fresh from the source (first iteration of the first align), 2 = re-seeded
    // (first iteration of the second align: the source moved by the first align's final transform, cache kept, bound
    // lowered by the distance between the two positions), 0 = as the previous iteration left it
    int lmode = stage == 0 ? 1 : 2;
    while (true) {
      float Mr[12];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        Mr[c * 3 + 0] = S.st.inc_T[c * 4 + 0];
        Mr[c * 3 + 1] = S.st.inc_T[c * 4 + 1];
        Mr[c * 3 + 2] = S.st.inc_T[c * 4 + 2];
      }
      const int apply = S.st.apply_inc;
      const int iter = S.st.iterations;
      const bool want_corr = A.corr_out != nullptr && stage == 0 && iter < A.corr_iters;
      int* CO = want_corr ? A.corr_out + ((size_t)iter * A.n_pairs_total + pair) * A.wstride : nullptr;
      double acc[16];
      int ncorr = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.0;
      bool any_b = false;  // a re-query pass ran after the phase-A sums were handed over (block-uniform)
      const long long tit0 = DBG ? clock64() : 0;
      P_TICK(5);

      for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int jbase = chunk * P_CHUNK, jend = min(nloc, jbase + P_CHUNK);
        // ---------------- load (all loads of the chunk are independent: one L2 round trip)
        if (lmode == 1) {
#pragma unroll
          for (int k = 0; k < P_REGP; ++k) {
            const int j = jbase + k * P_THREADS + tid;
            if (j < jend) {
              const float4 p = __ldg(&SRC[j * CL + crank]);
              px[k] = p.x;
              py[k] = p.y;
              pz[k] = p.z;
            } else {
              px[k] = py[k] = pz[k] = NAN;
            }
            lbv[k] = 0.f;
            slot[k] = 0u;
          }
        } else if (lmode == 2 || !resident) {
          if (!resident) {
#pragma unroll
            for (int k = 0; k < P_REGP; ++k) {
              const int j = jbase + k * P_THREADS + tid;
              if (j < jend) {
                const float4 p = W[j];
                px[k] = p.x;
                py[k] = p.y;
                pz[k] = p.z;
                slot[k] = __float_as_uint(p.w);
                lbv[k] = LB[j];
              } else {
                px[k] = py[k] = pz[k] = NAN;
                lbv[k] = 0.f;
                slot[k] = 0u;
              }
            }
          }
          if (lmode == 2) {
#pragma unroll
            for (int k = 0; k < P_REGP; ++k) {
              const int j = jbase + k * P_THREADS + tid;
              if (j < jend) {
                const float4 s = __ldg(&SRC[j * CL + crank]);
                float3 q = make_float3(s.x, s.y, s.z);
                if (finite3(s.x, s.y, s.z)) q = xform_point(S.T1, s.x, s.y, s.z);  // == transformPointCloud(src, final)
                if (finite3(q.x, q.y, q.z) && finite3(px[k], py[k], pz[k])) {
                  lbv[k] = lbv[k] - __fmaf_rn(sqrt_approx(dist2_l2simple(q.x, q.y, q.z, px[k], py[k], pz[k])), 1.00001f, 1e-9f);
                } else {
                  lbv[k] = 0.f;
                  slot[k] = 0u;
                }
                px[k] = q.x;
                py[k] = q.y;
                pz[k] = q.z;
              }
            }
          }
        }
        // ---------------- phase A: move, test the certified cache, accumulate; undecided points are flagged
        unsigned flags = 0u;
#pragma unroll
        for (int k = 0; k < P_REGP; ++k) {
          const int j = jbase + k * P_THREADS + tid;
          const int i = j * CL + crank;  // source index (correspondence dump)
          float x = px[k], y = py[k], z = pz[k];
          const bool fin = finite3(x, y, z);
          if (apply && fin) {
            const float qx = fadd(fadd(fadd(fmul(Mr[0], x), fmul(Mr[3], y)), fmul(Mr[6], z)), Mr[9]);
            const float qy = fadd(fadd(fadd(fmul(Mr[1], x), fmul(Mr[4], y)), fmul(Mr[7], z)), Mr[10]);
            const float qz = fadd(fadd(fadd(fmul(Mr[2], x), fmul(Mr[5], y)), fmul(Mr[8], z)), Mr[11]);
            // the bound decays by (an upper bound of) the distance this point just moved
            lbv[k] = lbv[k] - __fmaf_rn(sqrt_approx(dist2_l2simple(qx, qy, qz, x, y, z)), 1.00001f, 1e-9f);
            px[k] = x = qx;
            py[k] = y = qy;
            pz[k] = z = qz;
          }
          if (!fin) {
            if (want_corr && j < jend) CO[i] = -1;
            continue;
          }
          const int kp = (int)slot[k] - 1;
          bool valid;
          float bd = INFINITY;
          float qx = 0.f, qy = 0.f, qz = 0.f;
          if (kp >= 0) {
            qx = S.tx[kp];
            qy = S.ty[kp];
            qz = S.tz[kp];
            bd = dist2_l2simple(x, y, z, qx, qy, qz);
            valid = __fmaf_rn(sqrt_approx(bd), 1.0001f, 1e-7f) < lbv[k];
          } else {
            valid = lbv[k] > r;
          }
          if (!valid) {
            flags |= 1u << k;
            continue;
          }
          const bool ok = kp >= 0 && !((double)bd > prm.max_dist_sqr);
          if (want_corr) CO[i] = ok ? (int)TI[kp] : -1;
          if (ok) icp_accumulate16(acc, ncorr, x, y, z, qx, qy, qz, bd);
        }
        P_TICK(0);
        // ---------------- ordered compaction of the flagged points (thread-major, then k): block scan of the counts.
        // On the last chunk the warp totals of the sums collected so far travel with the same barrier: an iteration in
        // which the cache decides every point (the steady state) needs no further barrier before the solve.
        const bool last_chunk = chunk + 1 == nchunks;
        if (last_chunk) {
          const double v = warp_reduce16(acc, lane);
          const int n = __reduce_add_sync(0xffffffffu, ncorr);
          if ((lane & 1) == 0) S.red[0][wid][lane >> 1] = v;
          if (lane == 0) S.redn[0][wid] = n;
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[k] = 0.0;  // from here on: contributions of this chunk's re-queries
          ncorr = 0;
        }
        const int mycnt = __popc(flags);
        int incl = mycnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) S.wtot[wpar][wid] = incl;
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < P_WARPS; ++w) {
          const int c = S.wtot[wpar][w];
          woff += (w < wid) ? c : 0;
          total += c;
        }
        wpar ^= 1;
        const int myoff = woff + incl - mycnt;
        P_TICK(1);
        if (DBG && tid == 0) n_rescan += total;
        long long tb0 = 0;
        if (DBG && tid == 0) tb0 = clock64();
        // ---------------- phase B in rounds of P_WLCAP items
        for (int rbase = 0; rbase < total; rbase += P_WLCAP) {
          const int nw = min(P_WLCAP, total - rbase);
          {
            int pos = myoff - rbase;
#pragma unroll
            for (int k = 0; k < P_REGP; ++k) {
              if (flags & (1u << k)) {
                if (pos >= 0 && pos < P_WLCAP)
                  S.wl[pos] = make_float4(px[k], py[k], pz[k], __uint_as_float((slot[k] << 16) | (unsigned)(k * P_THREADS + tid)));
                ++pos;
              }
            }
          }
          __syncthreads();
          // lanes per item by the length of the list: a short list is bound by the dependent chain of one re-query
          if (nw <= P_WARPS)
            persist_requery_pass<32>(S, TI, nw, wid, lane, inv_cs, rmax, slack, ox, oy, oz, span, spany, spanz, prm.max_dist_sqr, CO, jbase, CL, crank, acc, ncorr);
          else if (nw <= 2 * P_WARPS)
            persist_requery_pass<16>(S, TI, nw, wid, lane, inv_cs, rmax, slack, ox, oy, oz, span, spany, spanz, prm.max_dist_sqr, CO, jbase, CL, crank, acc, ncorr);
          else if (nw <= 4 * P_WARPS)
            persist_requery_pass<8>(S, TI, nw, wid, lane, inv_cs, rmax, slack, ox, oy, oz, span, spany, spanz, prm.max_dist_sqr, CO, jbase, CL, crank, acc, ncorr);
          else
            persist_requery_pass<4>(S, TI, nw, wid, lane, inv_cs, rmax, slack, ox, oy, oz, span, spany, spanz, prm.max_dist_sqr, CO, jbase, CL, crank, acc, ncorr);
          __syncthreads();
          {  // owners take the new bound and cache slot back
            int pos = myoff - rbase;
#pragma unroll
            for (int k = 0; k < P_REGP; ++k) {
              if (flags & (1u << k)) {
                if (pos >= 0 && pos < P_WLCAP) {
                  const float4 e = S.wl[pos];
                  lbv[k] = e.x;
                  slot[k] = __float_as_uint(e.w);
                }
                ++pos;
              }
            }
          }
          if (rbase + P_WLCAP < total) __syncthreads();  // the work list is reused by the next round
        }
        if (last_chunk && total > 0) any_b = true;
        if (DBG && tid == 0 && A.dbg_iter && blockIdx.x == 0 && dbg_it < 256) {
          A.dbg_iter[3 * dbg_it + 0] += total;
          A.dbg_iter[3 * dbg_it + 1] += clock64() - tb0;
        }
        // ---------------- streamed slices go back to global memory
        if (!resident) {
#pragma unroll
          for (int k = 0; k < P_REGP; ++k) {
            const int j = jbase + k * P_THREADS + tid;
            if (j < jend) {
              W[j] = make_float4(px[k], py[k], pz[k], __uint_as_float(slot[k]));
              LB[j] = lbv[k];
            }
          }
        }
        P_TICK(2);
      }
      lmode = 0;
      // ---------------- the re-queried points' sums (second half of the reduction, only if there were any)
      if (any_b) {
        const double v = warp_reduce16(acc, lane);
        const int n = __reduce_add_sync(0xffffffffu, ncorr);
        if ((lane & 1) == 0) S.red[1][wid][lane >> 1] = v;
        if (lane == 0) S.redn[1][wid] = n;
        __syncthreads();
      }
      // ---------------- warp 0: CTA totals -> cluster exchange (st.async into every CTA's xch + its mbarrier) -> solve
      const int buf = parity;
      parity ^= 1;
      if (wid == 0) {
        double v = 0;
        if (lane == 0) {
          int n = 0;
#pragma unroll
          for (int w = 0; w < P_WARPS; ++w) n += S.redn[0][w];
          if (any_b) {
#pragma unroll
            for (int w = 0; w < P_WARPS; ++w) n += S.redn[1][w];
          }
          v = (double)n;
        } else if (lane < NRED) {
#pragma unroll
          for (int w = 0; w < P_WARPS; ++w) v += S.red[0][w][lane - 1];
          if (any_b) {
#pragma unroll
            for (int w = 0; w < P_WARPS; ++w) v += S.red[1][w][lane - 1];
          }
        }
        if (CL > 1) {
          // Buffer `buf` is rewritten two iterations later; a sibling can only be that far ahead after it has received
          // this CTA's partials of the NEXT iteration, i.e. after this CTA has finished reading the buffer.
          const unsigned mb = smem_u32(&S.mbar[buf]);
          if (lane < NRED) {
            const unsigned slot_addr = smem_u32(&S.xch[buf][crank][lane]);
            for (int rk = 0; rk < CL; ++rk) st_async_f64(mapa_u32(slot_addr, rk), v, mapa_u32(mb, rk));
          }
          if (lane == 0) mbar_arrive_expect_tx(mb, (unsigned)(CL * NRED * sizeof(double)));
          mbar_wait_cluster(mb, (mphase >> buf) & 1u);
          if (lane < NRED) {
            v = 0;
            for (int rk = 0; rk < CL; ++rk) v += S.xch[buf][rk][lane];  // fixed rank order: every CTA gets the same bits
          }
        }
        if (lane < NRED) S.tot[lane] = v;
        __syncwarp();
        P_TICK(3);
        icp_solve_pair(&S.st, S.tot, prm, &S.flags[1], lane);
        P_TICK(4);
      }
      mphase ^= 1u << buf;
      __syncthreads();
      if (DBG && tid == 0 && A.dbg_iter && blockIdx.x == 0 && dbg_it < 256) A.dbg_iter[3 * dbg_it + 2] = clock64() - tit0;
      if (DBG) ++dbg_it;
      if (S.st.done) break;
    }
    if (tid == 0 && crank == 0) (stage == 0 ? A.st1 : A.st2)[pair] = S.st;
    if (stage + 1 < A.n_stages) {
      // second align of the pair (icp:108-111): source = the first align's output (final applied to the ORIGINAL source,
      // as Registration::align writes it), guess = identity, a fresh convergence state -- the target replica and the
      // certified cache stay where they are
      __syncthreads();
      if (tid < 16) S.T1[tid] = S.st.final_T[tid];
      __syncthreads();
      if (tid == 0) {
        mat4_identity(S.st.final_T);
        mat4_identity(S.st.inc_T);
        S.st.prev_mse = A.prev2 ? A.prev2[pair] : DBL_MAX;
        S.st.mse = 0.0;
        S.st.iterations = 0;
        S.st.state = RSPCL_CONV_NOT_CONVERGED;
        S.st.converged = 0;
        S.st.done = 0;
        S.st.n_corr = 0;
        S.st.apply_inc = 0;
        S.flags[1] = 1;
      }
      __syncthreads();
    }
  }
  if (CL > 1) cg::this_cluster().sync();  // no CTA leaves while a sibling may still store into its shared memory
  if (DBG && A.dbg && tid == 0) {
    long long* D = A.dbg + ((size_t)pair * A.tidx_rep + crank) * 8;
    for (int k = 0; k < 6; ++k) D[k] = tph[k];
    D[6] = n_rescan;
    D[7] = clock64() - t_begin;
  }
#undef P_TICK
}
