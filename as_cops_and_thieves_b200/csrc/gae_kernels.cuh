// gae_kernels.cuh — MAPPO GAE as a parallel affine scan over T, and the in-place advantage normalisation
// (north-star item 4, the HBM-bound kernels).  Included by cat_b200.cu inside its anonymous namespace.
#pragma once
// ------------------------------------------------------------------ GAE
// adv_t = delta_t + c_t * adv_{t+1} with delta_t = r_t - V_t + gamma * nd_t * V_{t+1}, c_t = gamma * lambda * nd_t is a
// first-order linear recurrence, i.e. a scan of affine maps, so it parallelises over T as well as over
// the (world, agent) columns.  One CTA owns 32 columns (lanes: 128-B coalesced rows) and walks T backwards
// in chunks of kGaeSegs segments x kGaeS steps: warp s loads segment s of the chunk (all 8 x {r, V, done}
// loads issued up front), folds it into one affine map (A, B); the kGaeSegs maps of a column are combined
// by a shuffle scan (through shared memory, kGaeSegs lanes per column), and every thread then replays its
// steps from registers and writes advantages / returns.  Every input byte is read once and every output
// byte written once: 9 B read + 8 B written per sample.  CTAs are small (128 threads, 6 per SM) and the next
// chunk's loads are issued before the scan (software pipeline), so loads stay in flight during the scan and
// store phases.  sum / sum^2 of the advantages are reduced per
// CTA and added in fp64 for the normalisation pass.
#ifndef CAT_GAE_SEGS
#define CAT_GAE_SEGS 4   // measured on B200 (gpurun_out/prof_gae4.log): 4 segments x 32 columns, 6 CTAs per SM is the fastest shape
#endif
#ifndef CAT_GAE_MIN_CTAS
#define CAT_GAE_MIN_CTAS 6
#endif
constexpr int kGaeCols = 32, kGaeSegs = CAT_GAE_SEGS, kGaeS = 8;
constexpr int kGaeColsPerWarp = kGaeCols / kGaeSegs;  // scan phase: each warp scans 4 columns, 8 lanes per column

struct GaeChunk {          // one thread's 8 steps of one chunk, as loaded
  float r[kGaeS], v[kGaeS], vnext;
  uint32_t done;           // bit i: done flag of step i
};

// kFull: every step of the chunk exists (t >= 0) and every column of the CTA is < M, so nothing is clamped or
// predicated.  Index: element offsets are formed in 32 bits when T*M < 2^31 (one IMAD.WIDE per access instead of
// 64-bit multiply-adds — address arithmetic was a third of the kernel's instructions and the kernel is close to
// issue-bound once the loads are batched).
template <bool kFull, typename Index>
__device__ __forceinline__ void gae_load(GaeChunk& ck, const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                         const float* __restrict__ values, int base, int seg, int M, int col, bool valid,
                                         float carry_v) {
  // Branch-free: every address is clamped into the array so that all 24 loads issue back to back (a predicated
  // load per step would make the compiler wait for each step's `done` byte before issuing the next step's loads);
  // steps before t = 0 / columns beyond M are masked afterwards.
  const int ccol = kFull ? col : (valid ? col : 0);
  uint32_t raw[kGaeS];
#pragma unroll
  for (int i = 0; i < kGaeS; ++i) {
    const Index idx = (Index)(kFull ? base + i : max(base + i, 0)) * (Index)M + (Index)ccol;
    ck.r[i] = __ldcs(rewards + idx); ck.v[i] = __ldcs(values + idx); raw[i] = __ldcs(dones + idx);
  }
  ck.done = 0;
#pragma unroll
  for (int i = 0; i < kGaeS; ++i) ck.done |= (raw[i] ? 1u : 0u) << i;
  ck.vnext = carry_v;  // V_{t+1} of this thread's last step: the first V of the later segment (seg 0: patched by the caller)
  if (seg > 0) ck.vnext = __ldg(values + (Index)(kFull ? base + kGaeS : max(base + kGaeS, 0)) * (Index)M + (Index)ccol);
}

template <bool kFull>
__device__ __forceinline__ void gae_fold(GaeChunk& ck, int base, float gamma, float gl, float& A, float& B) {
  A = 1.f; B = 0.f;
#pragma unroll
  for (int i = kGaeS - 1; i >= 0; --i) {
    if (kFull || base + i >= 0) {
      const float nd = (ck.done >> i) & 1u ? 0.f : 1.f;
      const float vn = (i == kGaeS - 1) ? ck.vnext : ck.v[i + 1];
      ck.r[i] = ck.r[i] - ck.v[i] + gamma * nd * vn;  // delta_t
      B = fmaf(gl * nd, B, ck.r[i]);
      A *= gl * nd;
    }
  }
}

template <bool kFull, typename Index>
__device__ __forceinline__ void gae_store(const GaeChunk& ck, float adv, int base, int M, int col, bool valid, float gl,
                                          float* __restrict__ returns, float* __restrict__ advantages, float& p1, float& p2) {
#pragma unroll
  for (int i = kGaeS - 1; i >= 0; --i) {
    const int t = base + i;
    if (kFull || (valid && t >= 0)) {
      const float nd = (ck.done >> i) & 1u ? 0.f : 1.f;
      adv = fmaf(gl * nd, adv, ck.r[i]);
      const Index idx = (Index)t * (Index)M + (Index)col;
      advantages[idx] = adv;
      __stcs(returns + idx, adv + ck.v[i]);
      p1 += adv; p2 = fmaf(adv, adv, p2);
    }
  }
}

// Programmatic dependent launch (cat_b200.cu launches these kernels with programmatic stream serialisation): a CTA may
// become resident while the previous kernel in the stream is still draining; nothing that kernel wrote may be touched
// before pdl_wait() returns (it returns when the previous grid has completed and its writes are visible).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// The call's {sum(adv), sum(adv^2)} without a memset in front of the kernel and without a round trip at the end of a CTA.
// stats (include/cat_b200.h, CAT_GAE_STATS_DOUBLES doubles zeroed once by their owner) holds two sum slots, a 64-bit
// count of the calls made on this buffer and a 64-bit ticket:  { slot0[2], slot1[2], calls, ticket }.
//   call number c accumulates (fire-and-forget reductions) into slot (c + 1) & 1, which the call before left zero;
//   at its START every CTA reads c and then draws a ticket — the CTA that draws the last one knows every CTA of this
//   launch has read c, so it zeroes the other slot (the previous call's sums, whose consumers are earlier in the stream)
//   for the next call, resets the ticket and publishes calls = c + 1;
//   afterwards the sums of the latest call are in slot (calls & 1) — what cat_adv_normalize_kernel reads.
// Both round trips happen while the CTA's first tiles are in flight.
__device__ __forceinline__ int stats_begin(double* stats) {
  unsigned long long* u = reinterpret_cast<unsigned long long*>(stats);
  const unsigned long long c = *reinterpret_cast<volatile unsigned long long*>(u + 4);
  // (c >> 63 is zero: it makes the ticket depend on the value read, so the read is complete before the ticket is drawn)
  if (atomicAdd(u + 5, 1ull + (c >> 63)) == (unsigned long long)gridDim.x - 1ull) {
    const int z = (int)(c & 1ull);
    stats[2 * z] = 0.0; stats[2 * z + 1] = 0.0;
    u[5] = 0ull;
    u[4] = c + 1ull;
  }
  return (int)((c + 1ull) & 1ull);
}
__device__ __forceinline__ void stats_add(double* stats, int slot, double a, double b) {
  atomicAdd(&stats[2 * slot], a);
  atomicAdd(&stats[2 * slot + 1], b);
}

template <typename Index>
__global__ void __launch_bounds__(kGaeCols* kGaeSegs, CAT_GAE_MIN_CTAS)
    cat_gae_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones, const float* __restrict__ values,
                   const float* __restrict__ last_values, float* __restrict__ returns, float* __restrict__ advantages,
                   double* __restrict__ stats, int T, int M, float gamma, float lam) {
  __shared__ float sA[kGaeSegs][kGaeCols + 1], sB[kGaeSegs][kGaeCols + 1];
  __shared__ float sCarryAdv[kGaeCols], sCarryV[kGaeCols];
  __shared__ double sh1[kGaeSegs], sh2[kGaeSegs];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int col = blockIdx.x * kGaeCols + lane;
  const bool valid = col < M;
  pdl_wait();
  pdl_launch_dependents();
  int stats_slot = 0;
  if (threadIdx.x == 0) stats_slot = stats_begin(stats);
  const bool cols_full = blockIdx.x * kGaeCols + kGaeCols <= M;   // CTA-uniform
  const float gl = gamma * lam;
  constexpr int kChunk = kGaeSegs * kGaeS;
  float carry_adv = 0.f, carry_v = valid ? last_values[col] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  GaeChunk cur;
  if (cols_full && T >= kChunk) gae_load<true, Index>(cur, rewards, dones, values, T - (seg + 1) * kGaeS, seg, M, col, valid, carry_v);
  else gae_load<false, Index>(cur, rewards, dones, values, T - (seg + 1) * kGaeS, seg, M, col, valid, carry_v);
#pragma unroll 1
  for (int t_hi = T; t_hi > 0; t_hi -= kChunk) {
    const int base = t_hi - (seg + 1) * kGaeS;  // this thread's steps: base .. base + 7 (those >= 0)
    const bool full = cols_full && t_hi >= kChunk;
    if (seg == 0) cur.vnext = carry_v;            // known only now: V at the first step of the later chunk
    // software pipeline: the next (earlier) chunk's loads are in flight during this chunk's scan and stores
    GaeChunk nxt;
    if (t_hi > kChunk) {
      if (cols_full && t_hi >= 2 * kChunk) gae_load<true, Index>(nxt, rewards, dones, values, base - kChunk, seg, M, col, valid, 0.f);
      else gae_load<false, Index>(nxt, rewards, dones, values, base - kChunk, seg, M, col, valid, 0.f);
    }
    float A, B;
    if (full) gae_fold<true>(cur, base, gamma, gl, A, B); else gae_fold<false>(cur, base, gamma, gl, A, B);
    sA[seg][lane] = A; sB[seg][lane] = B;
    if (seg == kGaeSegs - 1) sCarryV[lane] = cur.v[0];  // V at the chunk's first step = V_{t+1} of the next chunk
    if (seg == 0) sCarryAdv[lane] = carry_adv;
    __syncthreads();
    {  // lane group g of warp `seg` scans column seg * 4 + g: sub-lane l holds the map of segment l
       // (x_{l+1} = B_l + A_l * x_l, x_0 = the advantage carried in from the later chunk)
      const int sl = lane & (kGaeSegs - 1), scol = seg * kGaeColsPerWarp + lane / kGaeSegs;
      float a = sA[sl][scol], b = sB[sl][scol];
#pragma unroll
      for (int d = 1; d < kGaeSegs; d <<= 1) {
        const float ap = __shfl_up_sync(0xFFFFFFFFu, a, d, kGaeSegs), bp = __shfl_up_sync(0xFFFFFFFFu, b, d, kGaeSegs);
        if (sl >= d) { b = fmaf(a, bp, b); a *= ap; }
      }
      const float x0 = sCarryAdv[scol];
      const float xout = fmaf(a, x0, b);                                // advantage at the first step of segment l
      float xin = __shfl_up_sync(0xFFFFFFFFu, xout, 1, kGaeSegs);       // = advantage entering segment l
      if (sl == 0) xin = x0;
      __syncwarp();
      sA[sl][scol] = xin;
      if (sl == kGaeSegs - 1) sB[0][scol] = xout;                       // advantage entering the next (earlier) chunk
    }
    __syncthreads();
    const float adv = sA[seg][lane];
    carry_adv = sB[0][lane];
    carry_v = sCarryV[lane];
    float p1 = 0.f, p2 = 0.f;
    if (full) gae_store<true, Index>(cur, adv, base, M, col, valid, gl, returns, advantages, p1, p2);
    else gae_store<false, Index>(cur, adv, base, M, col, valid, gl, returns, advantages, p1, p2);
    s1 += (double)p1; s2 += (double)p2;
    cur = nxt;
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
    s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
  }
  if (lane == 0) { sh1[seg] = s1; sh2[seg] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < kGaeSegs; ++i) { a += sh1[i]; b += sh2[i]; }
    stats_add(stats, stats_slot, a, b);
  }
}

// ---- the same scan, fed by TMA.  One CTA owns 64 columns; the chunk's 32 x 64 tiles of r, V (fp32) and done (u8)
// arrive in shared memory through three `cp.async.bulk.tensor.2d` copies (one per array, tensor maps encoded by
// the host per call) that complete on an mbarrier; a 3-stage ring keeps two chunks (36 KB) in flight per CTA while
// the third is folded, scanned and written — the in-flight bytes cost no registers, so four CTAs (1024 threads)
// fit per SM with 144 KB of loads outstanding.  Rows before t = 0 and columns beyond M are zero-filled by the TMA
// unit (out-of-bounds box) and masked in the arithmetic.  Needs M % 16 == 0 (tensor-map stride rule for the u8
// array); cat_gae falls back to the kernel above otherwise.
#ifndef CAT_GAE_TMA_STAGES
#define CAT_GAE_TMA_STAGES 3
#endif
#ifndef CAT_GAE_TMA_COLS
#define CAT_GAE_TMA_COLS 64
#endif
constexpr int kTmaCols = CAT_GAE_TMA_COLS, kTmaSegs = 4, kTmaRows = kTmaSegs * kGaeS, kTmaStages = CAT_GAE_TMA_STAGES;
constexpr int kTmaStageBytes = kTmaRows * kTmaCols * (4 + 4 + 1);   // 18432, a multiple of 128

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, unsigned long long* mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tmap), "r"(c0), "r"(c1),
      "r"((uint32_t)__cvta_generic_to_shared(mbar))
      : "memory");
}

template <typename Index>
__global__ void __launch_bounds__(kTmaCols* kTmaSegs, kTmaCols >= 64 ? 4 : 8)
    cat_gae_tma_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_v,
                       const __grid_constant__ CUtensorMap tm_d, const float* __restrict__ last_values,
                       float* __restrict__ returns, float* __restrict__ advantages, double* __restrict__ stats, int T, int M,
                       float gamma, float lam) {
  extern __shared__ __align__(128) unsigned char tiles[];
  __shared__ __align__(8) unsigned long long full[kTmaStages];
  __shared__ float sA[kTmaSegs][kTmaCols + 8], sB[kTmaSegs][kTmaCols + 8];   // +8: the scan's (segment, column) reads hit 32 banks
  __shared__ float sCarryAdv[kTmaCols], sCarryV[kTmaCols];
  __shared__ double sh1[8], sh2[8];
  const int tid = threadIdx.x, col = tid & (kTmaCols - 1), seg = tid / kTmaCols, lane = tid & 31;
  const int c0 = blockIdx.x * kTmaCols, gcol = c0 + col;
  const bool valid = gcol < M;
  const float gl = gamma * lam;
  const int nch = (T + kTmaRows - 1) / kTmaRows;
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {   // chunk k covers rows [T - (k+1)*32, T - k*32)
    const int s = k % kTmaStages, t_lo = T - (k + 1) * kTmaRows;
    unsigned char* base = tiles + s * kTmaStageBytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[s])),
                 "r"(kTmaStageBytes)
                 : "memory");
    tma_load_2d(base, &tm_r, c0, t_lo, &full[s]);
    tma_load_2d(base + kTmaRows * kTmaCols * 4, &tm_v, c0, t_lo, &full[s]);
    tma_load_2d(base + kTmaRows * kTmaCols * 8, &tm_d, c0, t_lo, &full[s]);
  };
  pdl_wait();                // everything above overlaps the tail of the previous kernel in the stream
  pdl_launch_dependents();   // ... and the next one may take this CTA's place the moment it exits
  int stats_slot = 0;
  if (tid == 0) {
    for (int k = 0; k < kTmaStages - 1 && k < nch; ++k) issue(k);
    stats_slot = stats_begin(stats);   // its two round trips overlap the tiles' flight
  }
  float carry_adv = 0.f, carry_v = valid ? last_values[gcol] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  const int rb = kTmaRows - (seg + 1) * kGaeS;   // this thread's 8 rows of the tile (segment 0 = the latest steps)
#pragma unroll 1
  for (int k = 0; k < nch; ++k) {
    const int s = k % kTmaStages, t_lo = T - (k + 1) * kTmaRows;
    {
      const uint32_t parity = (uint32_t)(k / kTmaStages) & 1u, addr = (uint32_t)__cvta_generic_to_shared(&full[s]);
      uint32_t done = 0;
#pragma unroll 1
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
    const float* R = reinterpret_cast<const float*>(tiles + s * kTmaStageBytes);
    const float* V = R + kTmaRows * kTmaCols;
    const unsigned char* D = reinterpret_cast<const unsigned char*>(V + kTmaRows * kTmaCols);
    float r[kGaeS], v[kGaeS];
    uint32_t dmask = 0;
#pragma unroll
    for (int i = 0; i < kGaeS; ++i) {
      r[i] = R[(rb + i) * kTmaCols + col]; v[i] = V[(rb + i) * kTmaCols + col];
      dmask |= (D[(rb + i) * kTmaCols + col] ? 1u : 0u) << i;
    }
    const float vnext = seg == 0 ? carry_v : V[(rb + kGaeS) * kTmaCols + col];
    float A = 1.f, B = 0.f;
#pragma unroll
    for (int i = kGaeS - 1; i >= 0; --i) {
      if (t_lo + rb + i >= 0) {
        const float nd = (dmask >> i) & 1u ? 0.f : 1.f;
        const float vn = (i == kGaeS - 1) ? vnext : v[i + 1];
        r[i] = r[i] - v[i] + gamma * nd * vn;
        B = fmaf(gl * nd, B, r[i]);
        A *= gl * nd;
      }
    }
    sA[seg][col] = A; sB[seg][col] = B;
    if (seg == kTmaSegs - 1) sCarryV[col] = v[0];
    if (seg == 0) sCarryAdv[col] = carry_adv;
    __syncthreads();
    {  // 4 lanes per column: sub-lane l holds the map of segment l (x_{l+1} = B_l + A_l * x_l)
      const int sl = tid & (kTmaSegs - 1), scol = tid / kTmaSegs;
      float a = sA[sl][scol], b = sB[sl][scol];
#pragma unroll
      for (int d = 1; d < kTmaSegs; d <<= 1) {
        const float ap = __shfl_up_sync(0xFFFFFFFFu, a, d, kTmaSegs), bp = __shfl_up_sync(0xFFFFFFFFu, b, d, kTmaSegs);
        if (sl >= d) { b = fmaf(a, bp, b); a *= ap; }
      }
      const float x0 = sCarryAdv[scol];
      const float xout = fmaf(a, x0, b);
      float xin = __shfl_up_sync(0xFFFFFFFFu, xout, 1, kTmaSegs);
      if (sl == 0) xin = x0;
      __syncwarp();
      sA[sl][scol] = xin;
      if (sl == kTmaSegs - 1) sB[0][scol] = xout;
    }
    __syncthreads();
    float adv = sA[seg][col];
    carry_adv = sB[0][col];
    carry_v = sCarryV[col];
    float p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int i = kGaeS - 1; i >= 0; --i) {
      const int t = t_lo + rb + i;
      if (valid && t >= 0) {
        const float nd = (dmask >> i) & 1u ? 0.f : 1.f;
        adv = fmaf(gl * nd, adv, r[i]);
        const Index idx = (Index)t * (Index)M + (Index)gcol;
        advantages[idx] = adv;
        __stcs(returns + idx, adv + v[i]);
        p1 += adv; p2 = fmaf(adv, adv, p2);
      }
    }
    s1 += (double)p1; s2 += (double)p2;
    __syncthreads();                                   // every thread has read stage s (and sA / sB): it can be refilled
    if (tid == 0 && k + kTmaStages - 1 < nch) issue(k + kTmaStages - 1);   // that stage was consumed in iteration k - 1
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
    s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
  }
  if (lane == 0) { sh1[tid >> 5] = s1; sh2[tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < (kTmaCols * kTmaSegs) / 32; ++i) { a += sh1[i]; b += sh2[i]; }
    stats_add(stats, stats_slot, a, b);
  }
}

// In-place (adv - mean) / (std + 1e-8): 4 B read + 4 B written per sample, four independent 16-B loads in
// flight per thread.
__global__ void __launch_bounds__(256) cat_adv_normalize_kernel(float* __restrict__ adv, long long n,
                                                                const double* __restrict__ stats, long long count) {
  // mean / 1/(std + eps) once per CTA (fp64 divide + sqrt), broadcast through shared memory
  __shared__ float s_fm, s_inv;
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    const int slot = (int)(reinterpret_cast<const unsigned long long*>(stats)[4] & 1ull);   // the latest call's sums
    const double sum = stats[2 * slot], sumsq = stats[2 * slot + 1];
    const double mean = sum / (double)count;
    double var = count > 1 ? (sumsq - (double)count * mean * mean) / (double)(count - 1) : 0.0;
    if (var < 0.0) var = 0.0;
    s_fm = (float)mean;
    s_inv = (float)(1.0 / (sqrt(var) + 1e-8));
  }
  __syncthreads();
  const float fm = s_fm, inv = s_inv;
  const long long n4 = n >> 2;
  float4* a4 = reinterpret_cast<float4*>(adv);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = a4[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u].x = (v[u].x - fm) * inv; v[u].y = (v[u].y - fm) * inv; v[u].z = (v[u].z - fm) * inv; v[u].w = (v[u].w - fm) * inv;
      a4[i + u * stride] = v[u];
    }
  }
  for (; i < n4; i += stride) {
    float4 v = a4[i];
    v.x = (v.x - fm) * inv; v.y = (v.y - fm) * inv; v.z = (v.z - fm) * inv; v.w = (v.w - fm) * inv;
    a4[i] = v;
  }
  for (long long j = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    adv[j] = (adv[j] - fm) * inv;
}

