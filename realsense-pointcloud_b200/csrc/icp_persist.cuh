// icp_persist.cuh -- persistent, shared-memory-resident ICP: one thread-block cluster per (source, target) pair runs
// EVERY iteration of align() inside a single launch.  (Included by icp.cu inside its anonymous namespace.)
//
// Why: edge clouds are ~10^4 points, so a per-iteration launch is bound by dependent L2 round trips and launch
// latency, not by HBM (profiles/r01_v1_summary.md).  B200 gives 227 KB of shared memory per CTA: the whole
// voxel-filtered target (<= 14336 points as SoA x/y/z = 168 KB; the 16-bit original indices stay in global memory) and its cell table (4096 x 8 B)
// fit in one SM, so every neighbour-cell probe and candidate read becomes an LDS (~30 cycles) instead of an L2 access
// (~300+ cycles), the convergence test never leaves the SM and the host never polls.
//   * cluster of CL CTAs per pair (CL = 4, 2 or 1 chosen from the batch size so the chip is filled): every CTA holds
//     a full replica of the target grid and owns 1/CL of the source points; the 17 fp64 partial sums are exchanged
//     through distributed shared memory (cluster.map_shared_rank) and each CTA redundantly runs the same solve, so no
//     broadcast is needed and all CTAs of a pair take the same convergence decision.
//   * the working source cloud stays in global memory (L2-resident, coalesced 16 B loads/stores, one independent
//     round trip per point per iteration) and is updated in place exactly like PCL's input_transformed.
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
constexpr int P_WL = 7168;                     // work-list entries (points per phase-A/B round)

// MUFU.SQRT (2^-22 relative error): only used for bounds that carry a 1e-5 safety margin
__device__ __forceinline__ float sqrt_approx(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

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

struct PersistSmem {
  float tx[P_NTMAX], ty[P_NTMAX], tz[P_NTMAX];
  uint2 tab[P_CAP];                 // {key, start << 16 | count}
  union {
    unsigned short fill[P_CAP];     // build-time cursors
    unsigned short wl[P_WL];        // per-iteration work list of points that need a cell scan (offset within the round)
  };
  int nwork;
  double red[P_WARPS][NRED];
  double part[2][NRED];             // this CTA's partial sums, double-buffered by iteration parity (read by the
                                    // other CTAs of the cluster through DSMEM)
  double tot[NRED];
  IcpState st;                      // replicated per CTA
  float M[16];
  int origin[3];                    // minimum cell coordinates of the target
  int cmax[3];
  int flags[4];                     // [0] build failed -> fallback, [1] dummy active counter
};

__device__ __forceinline__ unsigned p_hash(unsigned key) { return (key * 0x9E3779B1u) >> (32 - 12); }  // log2(P_CAP) = 12
static_assert(P_CAP == 4096, "p_hash assumes 4096 slots");

__device__ __forceinline__ int p_cell(float v, float inv_cs) { return __float2int_rd(__fmul_rn(v, inv_cs)); }

template <int CL>
__global__ void __launch_bounds__(P_THREADS, 1)
k_icp_persist(float4* __restrict__ work, const int* __restrict__ count, int wstride, IcpState* __restrict__ st_g,
              const float4* __restrict__ tgt, const int* __restrict__ tcount, int tstride, int shared_target,
              IcpDevParams prm, float inv_cs, int* __restrict__ first_corr, int* __restrict__ status,
              const int* __restrict__ order, float* __restrict__ lb, unsigned short* __restrict__ tidx_g,
              const unsigned short* __restrict__ carry_in, unsigned short* __restrict__ carry_out,
              long long* __restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char p_smem_raw[];
  PersistSmem& S = *reinterpret_cast<PersistSmem*>(p_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int pair = order[blockIdx.x / CL];  // largest pairs first: later waves fill the gaps (LPT)
  const int crank = (CL > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int tseg = shared_target ? 0 : pair;
  const int nt = tcount[tseg];
  const int ns = count[pair];
  const float4* T = tgt + (size_t)tseg * tstride;
  // original target index of every cell-sorted point: only needed to break exact distance ties and to report the first
  // correspondences, so it lives in global memory (one copy per CTA: the order inside a cell depends on each CTA's atomics) and the 2 B per
  // point it would cost in shared memory buy 2048 more resident target points
  unsigned short* TI = tidx_g + ((size_t)pair * CL + crank) * P_NTMAX;

  // ---------------------------------------------------------------- build the target replica in shared memory
  if (tid == 0) {
    S.origin[0] = S.origin[1] = S.origin[2] = INT_MAX;
    S.cmax[0] = S.cmax[1] = S.cmax[2] = INT_MIN;
    S.flags[0] = (nt > P_NTMAX) ? 1 : 0;
    S.flags[1] = 1;
    S.st = st_g[pair];
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
      for (int o = 16; o > 0; o >>= 1) {
        mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
        mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
      }
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
  // carried on): the final table is the same whatever order the atomics land in, so the cell-sorted target -- and with it
  // every cached slot number -- is reproducible from launch to launch (the fine align re-uses the coarse align's cache)
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
    const int total = __syncthreads_count(0) + 0;  // barrier
    (void)total;
    int* wtot = reinterpret_cast<int*>(&S.red[0][0]);
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(0xffffffffu, occ, o);
    if (lane == 0) wtot[wid] = occ;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < P_WARPS; ++w) t += wtot[w];
      if (t > (P_CAP * 85) / 100) S.flags[0] = 1;
    }
    __syncthreads();
  }
  if (S.flags[0]) {  // uniform across the CTA (and across the cluster: every CTA builds the same replica)
    if (tid == 0 && crank == 0) status[pair] = 1;
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
    int* wtot = reinterpret_cast<int*>(&S.red[0][0]);
    if (lane == 31) wtot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = lane < P_WARPS ? wtot[lane] : 0, vi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, vi, o);
        if (lane >= o) vi += t;
      }
      if (lane < P_WARPS) wtot[lane] = vi - v;
    }
    __syncthreads();
    int excl = wtot[wid] + incl - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      S.tab[tid * PER + k].y = ((unsigned)excl << 16) | (unsigned)loc[k];
      excl += loc[k];
    }
  }
  __syncthreads();
  // pass B: scatter the points into cell order
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
  __syncthreads();
  // Order every cell by original index (insertion sort, cells hold a few dozen points at most) so that the replica no
  // longer depends on the order in which the atomics of pass B happened to land: slot numbers then mean the same target
  // point in the next launch for the same target (cache carried from the coarse to the fine align).
  for (int sl = tid; sl < P_CAP; sl += P_THREADS) {
    const uint2 e = S.tab[sl];
    if (e.x == P_EMPTY) continue;
    const int b = (int)(e.y >> 16), n = (int)(e.y & 0xFFFFu);
    // (a cell with hundreds of points -- a gate of the size of the scene -- is left as it landed: one thread sorting it
    // would take O(n^2) global round trips; ties are still resolved by TI and carried slots are validated, so only the
    // reproducibility of those slot numbers is given up)
    if (n > 48) continue;
    for (int a = b + 1; a < b + n; ++a) {
      const unsigned short id = TI[a];
      const float x = S.tx[a], y = S.ty[a], z = S.tz[a];
      int c = a - 1;
      while (c >= b && TI[c] > id) {
        TI[c + 1] = TI[c];
        S.tx[c + 1] = S.tx[c];
        S.ty[c + 1] = S.ty[c];
        S.tz[c + 1] = S.tz[c];
        --c;
      }
      TI[c + 1] = id;
      S.tx[c + 1] = x;
      S.ty[c + 1] = y;
      S.tz[c + 1] = z;
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- iterations
  float4* W = work + (size_t)pair * wstride;
  float* LB = lb + (size_t)pair * wstride;
  const float r = prm.search_r;               // gate * 1.01
  const float rmax = __fdividef(0.485f, inv_cs);  // largest ball that touches <= 2 cells per axis (~1.99 x gate)
  const float slack = 0.5f * r;
  const int span = S.cmax[0] - ox, spany = S.cmax[1] - oy, spanz = S.cmax[2] - oz;
  const int share = (ns + CL - 1) / CL;  // this CTA's contiguous slice [lo, hi) of the source
  const int lo = crank * share, hi = min(ns, lo + share);
  if (carry_in) {
    // The working cloud arrives with the cache of a previous align of the same pair (coarse -> fine): a cached slot is
    // only trusted if it still names the same target point in THIS launch's replica (original index carried alongside);
    // anything else starts uncached.  The bounds were already lowered by the distance each point moved in between.
    const unsigned short* CI = carry_in + (size_t)pair * wstride;
    for (int i = lo + tid; i < hi; i += P_THREADS) {
      const unsigned wb = __float_as_uint(W[i].w);
      const int kp = (int)(wb >> 16) - 1;
      if (kp >= 0 && (kp >= P_NTMAX || TI[kp] != CI[i])) {
        W[i].w = __uint_as_float(wb & 0xFFFFu);
        LB[i] = 0.f;
      }
    }
    __syncthreads();
  }
  int parity = 0;
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tc = clock64(), n_rescan = 0;  // debug phase timers (thread 0; only stored if dbg != nullptr)
#define P_TICK(k) do { if (dbg && tid == 0) { const long long n_ = clock64(); tph[k] += n_ - tc; tc = n_; } } while (0)
  if (dbg && tid == 0) dbg[((size_t)pair * CL + crank) * 8 + 6] = tc;
  while (true) {
    if (tid < 16) S.M[tid] = S.st.inc_T[tid];
    __syncthreads();
    P_TICK(5);
    const int apply = S.st.apply_inc;
    const bool want_corr = first_corr != nullptr && S.st.iterations == 0;
    double acc[NRED];
#pragma unroll
    for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
    // Certified cache (exact).  LB[i] is a lower bound of the true distance from source point i to EVERY target point
    // other than its cached match (to every target point if none is cached).  While the cached match is closer than
    // that bound it is the strict nearest neighbour; while the bound exceeds the gate an unmatched point stays
    // unmatched: no cell is visited (phase A, straight-line code, two points in flight per thread).  The few points
    // whose bound no longer decides go to a shared-memory work list and are rescanned load-balanced over the whole CTA
    // (phase B, <= 2x2x2 cells), which also renews their bound.
    for (int base = lo; base < hi; base += P_WL) {  // one round unless the slice exceeds the work list
      const int end = min(hi, base + P_WL);
      if (tid == 0) S.nwork = 0;
      __syncthreads();
      // ---------------- phase A (software-pipelined: the next point's loads are in flight while this one is processed)
      {
        int i = base + tid;
        float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
        float ln = 0.f;
        if (i < end) {
          pn = W[i];
          ln = LB[i];
        }
        for (; i < end; i += P_THREADS) {
          float4 p = pn;
          float lbv = ln;
          if (i + P_THREADS < end) {
            pn = W[i + P_THREADS];
            ln = LB[i + P_THREADS];
          }
          const bool fin = finite3(p.x, p.y, p.z);
          if (apply && fin) {
            const float3 q = xform_point(S.M, p.x, p.y, p.z);
            // the bound decays by (an upper bound of) the distance this point just moved
            lbv = lbv - __fmaf_rn(sqrt_approx(dist2_l2simple(q.x, q.y, q.z, p.x, p.y, p.z)), 1.00001f, 1e-9f);
            p.x = q.x;
            p.y = q.y;
            p.z = q.z;
            W[i] = p;
          }
          LB[i] = lbv;
          if (!fin) {
            if (want_corr) first_corr[(size_t)pair * wstride + (int)(__float_as_uint(p.w) & 0xFFFFu)] = -1;
            continue;
          }
          // .w = (cached match slot + 1) << 16 | original source index
          const unsigned wbits = __float_as_uint(p.w);
          const int kp = (int)(wbits >> 16) - 1;
          bool valid;
          float bd = INFINITY;
          if (kp >= 0) {
            bd = dist2_l2simple(p.x, p.y, p.z, S.tx[kp], S.ty[kp], S.tz[kp]);
            valid = __fmaf_rn(sqrt_approx(bd), 1.0001f, 1e-7f) < lbv;
          } else {
            valid = lbv > r;
          }
          if (!valid) {
            S.wl[atomicAdd(&S.nwork, 1)] = (unsigned short)(i - base);
            continue;
          }
          const bool ok = kp >= 0 && !((double)bd > prm.max_dist_sqr);
          if (want_corr) first_corr[(size_t)pair * wstride + (int)(wbits & 0xFFFFu)] = ok ? (int)TI[kp] : -1;
          if (ok) icp_accumulate(acc, p.x, p.y, p.z, S.tx[kp], S.ty[kp], S.tz[kp], bd);
        }
      }
      P_TICK(0);
      __syncthreads();
      P_TICK(1);
      if (dbg && tid == 0) n_rescan += S.nwork;
      // ---------------- phase B: P_G lanes per work item (each probes 8 / P_G neighbour cells), top-2 merged with xor shuffles
      const int nwork = S.nwork;
      for (int jb = wid * (32 / P_G); jb < nwork; jb += P_WARPS * (32 / P_G)) {  // warp-uniform trip count (full-mask shuffles)
        const int j = jb + lane / P_G, c = lane % P_G;
        const bool active = j < nwork;
        const int i = base + (active ? (int)S.wl[j] : 0);
        const float4 p = W[i];  // already transformed by phase A
        const unsigned wbits = __float_as_uint(p.w);
        const int orig = (int)(wbits & 0xFFFFu);
        int kp = (int)(wbits >> 16) - 1;
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
        unsigned cell[8 / P_G];                // this lane's neighbour cells: start << 16 | count (0: empty / duplicate / outside)
        {
          const int x0 = p_cell(p.x - rr, inv_cs) - ox, x1 = p_cell(p.x + rr, inv_cs) - ox;
          const int y0 = p_cell(p.y - rr, inv_cs) - oy, y1 = p_cell(p.y + rr, inv_cs) - oy;
          const int z0 = p_cell(p.z - rr, inv_cs) - oz, z1 = p_cell(p.z + rr, inv_cs) - oz;
#pragma unroll
          for (int q = 0; q < 8 / P_G; ++q) {
            const int cc = c + q * P_G;
            cell[q] = 0u;
            const bool dup = ((cc & 1) && x1 == x0) || ((cc & 2) && y1 == y0) || ((cc & 4) && z1 == z0);
            const int ix = (cc & 1) ? x1 : x0, iy = (cc & 2) ? y1 : y0, iz = (cc & 4) ? z1 : z0;
            const bool inside = (unsigned)ix <= (unsigned)span && (unsigned)iy <= (unsigned)spany && (unsigned)iz <= (unsigned)spanz;
            if (active && !dup && inside) {
              const unsigned key = ((unsigned)ix << 20) | ((unsigned)iy << 10) | (unsigned)iz;
              unsigned s = p_hash(key);
              uint2 e = S.tab[s];
              while (e.x != key && e.x != P_EMPTY) {
                s = (s + 1) & (P_CAP - 1);
                e = S.tab[s];
              }
              if (e.x == key) cell[q] = e.y;
            }
          }
        }
        // The lanes of the group now walk the (<= 8) occupied cells together, lane `c` taking every P_G-th candidate:
        // consecutive lanes read consecutive shared-memory words (no bank conflicts inside a group) and the work is
        // balanced however unevenly the points are spread over the cells.
#pragma unroll
        for (int q = 0; q < 8 / P_G; ++q) {
#pragma unroll 1
          for (int sl = 0; sl < P_G; ++sl) {
            const unsigned ey = __shfl_sync(0xffffffffu, cell[q], sl, P_G);
            const int b = (int)(ey >> 16), en = b + (int)(ey & 0xFFFFu);
            for (int k = b + c; k < en; k += P_G) {
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
          }
        }
#pragma unroll
        for (int o = 1; o < P_G; o <<= 1) {  // top-2 merge across the group (every lane ends with the group result)
          const float obd = __shfl_xor_sync(0xffffffffu, bd, o), o2 = __shfl_xor_sync(0xffffffffu, d2nd, o);
          const int ok_ = __shfl_xor_sync(0xffffffffu, kbest, o);
          if (obd < bd || (obd == bd && ok_ >= 0 && kbest >= 0 && TI[ok_] < TI[kbest])) {
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
          LB[i] = fminf(sqrt_approx(d2nd), edge) * 0.9999f;
          W[i].w = __uint_as_float(((unsigned)(kbest + 1) << 16) | (unsigned)orig);
          const bool ok = kbest >= 0 && !((double)bd > prm.max_dist_sqr);
          if (want_corr) first_corr[(size_t)pair * wstride + orig] = ok ? (int)TI[kbest] : -1;
          if (ok) icp_accumulate(acc, p.x, p.y, p.z, S.tx[kbest], S.ty[kbest], S.tz[kbest], bd);
        }
      }
      if (base + P_WL < hi) __syncthreads();  // the work list is reused by the next round
    }
    P_TICK(2);
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      const double v = warp_sum(acc[k]);
      if (lane == 0) S.red[wid][k] = v;
    }
    __syncthreads();
    const int buf = parity;
    parity ^= 1;
    if (tid < NRED) {
      double v = 0;
#pragma unroll
      for (int w = 0; w < P_WARPS; ++w) v += S.red[w][tid];
      S.part[buf][tid] = v;
    }
    if (CL > 1) {
      // One cluster barrier per iteration: buffer `buf` is rewritten two iterations later, after every CTA has passed
      // the next barrier, i.e. after every CTA has finished reading it.
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();
      if (tid < NRED) {
        double v = 0;
        for (int rk = 0; rk < CL; ++rk) v += *cluster.map_shared_rank(&S.part[buf][tid], rk);  // fixed rank order
        S.tot[tid] = v;
      }
      __syncthreads();
    } else {
      __syncthreads();
      if (tid < NRED) S.tot[tid] = S.part[buf][tid];
      __syncthreads();
    }
    P_TICK(3);
    if (wid == 0) icp_solve_pair(&S.st, S.tot, prm, &S.flags[1], lane);
    P_TICK(4);
    __syncthreads();
    if (S.st.done) break;
  }
  if (carry_out) {  // what the next align of this pair needs to trust the cache: the target point behind every slot
    unsigned short* CO = carry_out + (size_t)pair * wstride;
    for (int i = lo + tid; i < hi; i += P_THREADS) {
      const int kp = (int)(__float_as_uint(W[i].w) >> 16) - 1;
      CO[i] = kp >= 0 ? TI[kp] : (unsigned short)0xFFFFu;
    }
  }
  if (CL > 1) cg::this_cluster().sync();  // no CTA leaves while a sibling may still read its partials
  if (tid == 0 && crank == 0) st_g[pair] = S.st;
  if (dbg && tid == 0) {
    long long* D = dbg + ((size_t)pair * CL + crank) * 8;
    for (int k = 0; k < 6; ++k) D[k] = tph[k];
    D[7] = clock64() - D[6];
    D[6] = n_rescan;
  }
#undef P_TICK
}
