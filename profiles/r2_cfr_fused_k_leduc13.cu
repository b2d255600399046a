struct FusedParams {
const int * mem_opp0;
const int * mem_opp1;
const int * mem_pat0;
const int * mem_pat1;
const int * mem_ridx0;
const int * mem_ridx1;
const int * grp_ridx0;
const int * grp_ridx1;
const int * mem_pidx0;
const int * mem_pidx1;
const double * mem_pc0;
const double * mem_pc1;
const int * grp_fid0;
const int * grp_fid1;
const int * grp_col0;
const int * grp_col1;
const int * grp_iset0;
const int * grp_iset1;
const double * paytab;
const signed char * paytab8;
const int * tg_tile0;
const int * tg_tile1;
const int * tg_fid0;
const int * tg_fid1;
const int * tg_col0;
const int * tg_col1;
const int * tg_iset0;
const int * tg_iset1;
const int * u_fid;
const double * u_pcs;
const double * u_cprob;
const double * u_pay1;
const int * fan_desc;
const double * fan_cp;
double * R;
double * S;
double * sigma_row;
double * sigF0;
double * sigF1;
double * portal_p1;
double * portal_p2;
double * pvl0;
double * pvl1;
double * pvA0;
double * pvA1;
double * pvA2;
double * pvA3;
double * pvA4;
double * pvA5;
double * pvA6;
double * pvA7;
double * pvB0;
double * pvB1;
double * pvB2;
double * pvB3;
double * pvB4;
double * pvB5;
double * pvB6;
double * pvB7;
unsigned long long * xfp0;
unsigned long long * xfp1;
unsigned long long * xfp2;
unsigned long long * xfp3;
unsigned long long * xfp4;
unsigned long long * xfp5;
unsigned long long * xfp6;
unsigned long long * xfp7;
unsigned long long * xfl;
long long nranks;
long long rank;
long long xbase;
unsigned char * flags;
unsigned int * bar;
unsigned long long * tl;
const double * br_sig;
double * br_reach1;
double * br_reach2;
double * br_pv1;
double * br_pv2;
double * br_top1;
double * br_top2;
double * br_out;
int * br_astar;
const int * tile_root;
const int * g_first_child;
const signed char * g_player;
long long n_top;
long long n_tl;
long long tl0;
long long tl1;
long long tl2;
long long tl3;
long long tl4;
long long tl5;
long long tl6;
long long tl7;
long long tl8;
long long num_isets;
long long first_group0;
long long end_group0;
long long first_group1;
long long end_group1;
long long tl_cap;
};
#define BLOCK 256
#define TEAM 256
#define SMS 257
#define TEAM_SYNC() __syncthreads()

__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifdef XDEBUG   // RLP_P2P_DEBUG: the last arriver of the cross-GPU barrier adds up where its time goes (bar[8..17], ns)
__device__ __forceinline__ void xdbg(unsigned int* bar, int i, unsigned long long& t0) {
  unsigned long long t = now_ns();
  atomicAdd((unsigned long long*)(bar + 8) + i, i ? t - t0 : 1ull);
  t0 = t;
}
#define XDBG(i) xdbg(bar, i, xdbg_t);
#else
#define XDBG(i)
#endif
// Grid barrier for co-resident CTAs (cooperative launch).  bar[0] counts arrivals (monotone over the launch); the LAST
// arriver of barrier number `k` writes k into a private flag word of every CTA (bar[32 + 32 * cta], 128 bytes apart,
// so spread over the L2 slices) and each waiting CTA polls only its own word: no address is polled by more than one
// thread.  (Polling bar[0] from ~300 CTAs saturates its L2 slice and slows every load that maps to that slice.)
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int k, unsigned int nb) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(bar, 1u);
    s_last = prev + 1u == k * nb;
    __threadfence();
  }
  __syncthreads();
  if (s_last) {
    for (unsigned int c = threadIdx.x; c < nb; c += BLOCK)
      asm volatile("st.relaxed.gpu.u32 [%0], %1;" :: "l"(bar + 32 + 32 * c), "r"(k) : "memory");
  } else if (threadIdx.x == 0) {
    const unsigned int* mine = bar + 32 + 32 * blockIdx.x;
    unsigned int v;
    for (;;) {
      asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= k) break;
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}
// The same barrier with the exchange step of the multi-GPU solve folded in: the subtree values were stored straight
// into every peer's buffer (NVLink peer stores) during the micro phase; the last CTA of this GPU to arrive publishes
// `token` in every peer's flag array and waits until every peer has published it here, then releases the local CTAs.
// System-scope fences order the peer stores of all local CTAs before the flag.  The wait is bounded: a lost peer
// sets bar[2] (the host reports it) instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier_x(const FusedParams* __restrict__ P, unsigned int k, unsigned int nb, unsigned long long token,
                                               bool pushed) {
  __shared__ int s_last;
  unsigned int* bar = P->bar;
  __syncthreads();
  if (threadIdx.x == 0) {
    // only the CTAs that stored into peer memory pay for a system-scope fence (hundreds of them at once are slow)
    if (pushed) __threadfence_system(); else __threadfence();
    const unsigned int prev = atomicAdd(bar, 1u);
    s_last = prev + 1u == k * nb;
    __threadfence();
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) {
      const int nr = (int)P->nranks, me = (int)P->rank;
      unsigned long long* const* peers = &P->xfp0;
      // ONE system-scope fence (it is what costs: it waits for this GPU's outstanding peer writes), then relaxed flag stores;
      // a st.release per peer would repeat that fence for every peer
      unsigned long long xdbg_t = 0; (void)xdbg_t;
      XDBG(0)
      __threadfence_system();
      XDBG(1)
      for (int r = 0; r < nr; ++r)
        if (r != me) asm volatile("st.relaxed.sys.u64 [%0], %1;" :: "l"(peers[r] + me), "l"(token) : "memory");
      XDBG(2)
      for (int r = 0; r < nr; ++r) {
        if (r == me) continue;
        unsigned long long v = 0;
        long long spins = 0;
        for (;;) {
          asm volatile("ld.relaxed.sys.u64 %0, [%1];" : "=l"(v) : "l"(P->xfl + r) : "memory");
          if (v >= token) break;
          if (++spins > (1ll << 25)) { bar[2] = 1u; break; }      // ~seconds: give up, flag the error
          __nanosleep(40);
        }
      }
      XDBG(3)
      __threadfence_system();
      XDBG(4)
    }
    __syncthreads();
    for (unsigned int c = threadIdx.x; c < nb; c += BLOCK)
      asm volatile("st.relaxed.gpu.u32 [%0], %1;" :: "l"(bar + 32 + 32 * c), "r"(k) : "memory");
  } else if (threadIdx.x == 0) {
    const unsigned int* mine = bar + 32 + 32 * blockIdx.x;
    unsigned int v;
    for (;;) {
      asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= k) break;
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}
// compute_regret_matching + normalise_probs (cfr_util.py:14-58) for action `a` of a set whose regrets are r[0..na)
__device__ __forceinline__ double rm_one(const double* r, int na, int a) {
  double mx = r[0];
  for (int b = 1; b < na; ++b) mx = r[b] > mx ? r[b] : mx;
  if (mx <= 0.0) return 1.0 / (double)na;
  double s = 0.0, c = 0.0, mine = 0.0;
  for (int b = 0; b < na; ++b) {
    double p = r[b] > 0.0 ? r[b] : 0.0;
    p = p > 1e-7 ? p : 1e-7;
    double t = s + p;
    if (fabs(s) >= fabs(p)) c += (s - t) + p; else c += (p - t) + s;
    s = t;
    if (b == a) mine = p;
  }
  return mine / (s + c);
}
__constant__ double FCP[9] = {0.041666666666666664,0.041666666666666664,0.041666666666666664,0.041666666666666664,0.041666666666666664,0.041666666666666664,0.041666666666666664,0.041666666666666664,0.041666666666666664};
struct MIds { int pat, ri, ro, pi, acol, afid; double pcv; int f[8]; };
__constant__ int MCJ0[11] = {0,0,1,1,1,2,2,2,3,3,3};
__constant__ int MCA0[11] = {0,1,0,1,2,0,1,2,0,1,2};
__constant__ int MCOFF0[4] = {0,2,5,8};
__constant__ int MNA0[4] = {2,3,3,3};
__constant__ int UCJ0[11] = {0,0,1,1,1,2,2,2,3,3,3};
__constant__ int UCA0[11] = {0,1,0,1,2,0,1,2,0,1,2};
__constant__ int UCOFF0[4] = {0,2,5,8};
__constant__ int UNA0[4] = {2,3,3,3};
__device__ __forceinline__ MIds micro_ids0(const FusedParams* __restrict__ P, long long g, int k, bool ok,
    long long ag, int aj, bool acc) {
  MIds id;
  const long long slot = ok ? g * 24 + k : 0;
  const long long gg = ok ? g : 0;
  id.pat = P->mem_pat0[slot]; id.ri = P->mem_ridx0[slot]; id.ro = P->grp_ridx0[gg];
  id.pcv = P->mem_pc0[slot];
  id.pi = P->mem_pidx0[slot];
  id.f[0] = P->mem_opp0[0LL * 140400LL + slot];
  id.f[1] = P->grp_fid0[gg * 4 + 0];
  id.f[2] = P->grp_fid0[gg * 4 + 1];
  id.f[3] = P->mem_opp0[1LL * 140400LL + slot];
  id.f[4] = P->mem_opp0[2LL * 140400LL + slot];
  id.f[5] = P->grp_fid0[gg * 4 + 2];
  id.f[6] = P->grp_fid0[gg * 4 + 3];
  id.f[7] = P->mem_opp0[3LL * 140400LL + slot];
  const long long agg = acc ? ag : 0;
  id.acol = P->grp_col0[agg * 4 + aj];
  id.afid = P->grp_fid0[agg * 4 + aj];
  return id;
}
__device__ __forceinline__ void micro_eval0(const FusedParams* __restrict__ P, const MIds& id, int tix, double w,
    const double* sig_in, double* sm_r, double* sm_o, int par) {
  const int pat = id.pat, ri = id.ri, ro = id.ro;
  const double pcv = id.pcv;
  const int pi = id.pi;
  const int f0 = id.f[0];
  const int f1 = id.f[1];
  const int f2 = id.f[2];
  const int f3 = id.f[3];
  const int f4 = id.f[4];
  const int f5 = id.f[5];
  const int f6 = id.f[6];
  const int f7 = id.f[7];
  const int4* pp = reinterpret_cast<const int4*>(P->paytab8 + (long long)pat * 16);
  const int4 pw0 = __ldg(pp + 0);
  const double x3 = (double)(int)(signed char)(pw0.x >> 0);
  const double x5 = (double)(int)(signed char)(pw0.x >> 8);
  const double x6 = (double)(int)(signed char)(pw0.x >> 16);
  const double x8 = (double)(int)(signed char)(pw0.x >> 24);
  const double x9 = (double)(int)(signed char)(pw0.y >> 0);
  const double x11 = (double)(int)(signed char)(pw0.y >> 8);
  const double x12 = (double)(int)(signed char)(pw0.y >> 16);
  const double x14 = (double)(int)(signed char)(pw0.y >> 24);
  const double x15 = (double)(int)(signed char)(pw0.z >> 0);
  const double x17 = (double)(int)(signed char)(pw0.z >> 8);
  const double x18 = (double)(int)(signed char)(pw0.z >> 16);
  const double x19 = (double)(int)(signed char)(pw0.z >> 24);
  const double x20 = (double)(int)(signed char)(pw0.w >> 0);
  const double x21 = (double)(int)(signed char)(pw0.w >> 8);
  const double x22 = (double)(int)(signed char)(pw0.w >> 16);
  const double q1_0 = __ldcg(P->portal_p1 + ro), q2_0 = __ldcg(P->portal_p2 + ri);
  const double s1 = __ldcg(sig_in + 0LL * 47008LL + f0);
  const double s2 = __ldcg(sig_in + 1LL * 47008LL + f0);
  const double s3 = __ldcg(sig_in + 0LL * 47008LL + f1);
  const double s4 = __ldcg(sig_in + 1LL * 47008LL + f1);
  const double s5 = __ldcg(sig_in + 0LL * 47008LL + f2);
  const double s6 = __ldcg(sig_in + 1LL * 47008LL + f2);
  const double s7 = __ldcg(sig_in + 2LL * 47008LL + f2);
  const double s8 = __ldcg(sig_in + 0LL * 47008LL + f3);
  const double s9 = __ldcg(sig_in + 1LL * 47008LL + f3);
  const double s10 = __ldcg(sig_in + 2LL * 47008LL + f3);
  const double s11 = __ldcg(sig_in + 0LL * 47008LL + f4);
  const double s12 = __ldcg(sig_in + 1LL * 47008LL + f4);
  const double s13 = __ldcg(sig_in + 2LL * 47008LL + f4);
  const double s14 = __ldcg(sig_in + 0LL * 47008LL + f5);
  const double s15 = __ldcg(sig_in + 1LL * 47008LL + f5);
  const double s16 = __ldcg(sig_in + 2LL * 47008LL + f5);
  const double s17 = __ldcg(sig_in + 0LL * 47008LL + f6);
  const double s18 = __ldcg(sig_in + 1LL * 47008LL + f6);
  const double s19 = __ldcg(sig_in + 2LL * 47008LL + f6);
  const double s20 = __ldcg(sig_in + 0LL * 47008LL + f7);
  const double s21 = __ldcg(sig_in + 1LL * 47008LL + f7);
  const double s22 = __ldcg(sig_in + 2LL * 47008LL + f7);
  const double q1_1 = q1_0;
  const double q2_1 = s1 * q2_0;
  const double q1_2 = q1_0;
  const double q2_2 = s2 * q2_0;
  const double q1_3 = s4 * q1_1;
  const double q2_3 = q2_1;
  const double q1_4 = s7 * q1_2;
  const double q2_4 = q2_2;
  const double q1_5 = q1_3;
  const double q2_5 = s10 * q2_3;
  const double q1_6 = q1_4;
  const double q2_6 = s13 * q2_4;
  const double q1_7 = s16 * q1_5;
  const double q2_7 = q2_5;
  double a16 = 0.0;
  a16 = a16 + s20 * x20;
  a16 = a16 + s21 * x21;
  a16 = a16 + s22 * x22;
  const double x16 = a16;
  double a13 = 0.0;
  a13 = a13 + s17 * x17;
  a13 = a13 + s18 * x18;
  a13 = a13 + s19 * x19;
  const double x13 = a13;
  {
    const double opp = pcv * q2_6;
    const double own = q1_6;
    const double vp = x13;
    sm_r[8 * SMS + tix] = w * (x17 - vp) * opp;
    sm_r[9 * SMS + tix] = w * (x18 - vp) * opp;
    sm_r[10 * SMS + tix] = w * (x19 - vp) * opp;
    sm_o[3 * SMS + tix] = w * own;
  }
  double a10 = 0.0;
  a10 = a10 + s14 * x14;
  a10 = a10 + s15 * x15;
  a10 = a10 + s16 * x16;
  const double x10 = a10;
  {
    const double opp = pcv * q2_5;
    const double own = q1_5;
    const double vp = x10;
    sm_r[5 * SMS + tix] = w * (x14 - vp) * opp;
    sm_r[6 * SMS + tix] = w * (x15 - vp) * opp;
    sm_r[7 * SMS + tix] = w * (x16 - vp) * opp;
    sm_o[2 * SMS + tix] = w * own;
  }
  double a7 = 0.0;
  a7 = a7 + s11 * x11;
  a7 = a7 + s12 * x12;
  a7 = a7 + s13 * x13;
  const double x7 = a7;
  double a4 = 0.0;
  a4 = a4 + s8 * x8;
  a4 = a4 + s9 * x9;
  a4 = a4 + s10 * x10;
  const double x4 = a4;
  double a2 = 0.0;
  a2 = a2 + s5 * x5;
  a2 = a2 + s6 * x6;
  a2 = a2 + s7 * x7;
  const double x2 = a2;
  {
    const double opp = pcv * q2_2;
    const double own = q1_2;
    const double vp = x2;
    sm_r[2 * SMS + tix] = w * (x5 - vp) * opp;
    sm_r[3 * SMS + tix] = w * (x6 - vp) * opp;
    sm_r[4 * SMS + tix] = w * (x7 - vp) * opp;
    sm_o[1 * SMS + tix] = w * own;
  }
  double a1 = 0.0;
  a1 = a1 + s3 * x3;
  a1 = a1 + s4 * x4;
  const double x1 = a1;
  {
    const double opp = pcv * q2_1;
    const double own = q1_1;
    const double vp = x1;
    sm_r[0 * SMS + tix] = w * (x3 - vp) * opp;
    sm_r[1 * SMS + tix] = w * (x4 - vp) * opp;
    sm_o[0 * SMS + tix] = w * own;
  }
  double a0 = 0.0;
  a0 = a0 + s1 * x1;
  a0 = a0 + s2 * x2;
  const double x0 = a0;
  { double* const* pvp = par ? &P->pvB0 : &P->pvA0; const int nr = (int)P->nranks;
    for (int r = 0; r < nr; ++r) pvp[r][pi] = x0; }
}
struct UIn0 {
  double pc0;
  double s1;
  double s2;
  double s3;
  double pc4;
  double s4;
  double s5;
  double x6;
  double s6;
  double pc7;
  double s7;
  double s8;
  double x9;
  double s9;
  double s10;
  double s11;
  double x12;
  double s12;
  double s13;
  double s14;
  double x15;
  double s15;
  double pc16;
  double s16;
  double s17;
  double x18;
  double s18;
  double s19;
  double s20;
  double x21;
  double s21;
  double s22;
  double pad_;
};
__device__ __forceinline__ void upper_load0(const FusedParams* __restrict__ P, long long i, int m, const int* sm_ut, const double* sig_in, const double* blk, UIn0& in) {
  const int f0 = sm_ut[25 + m];
  in.pc0 = P->u_pcs[0LL * 650LL + i];
  const int f1 = sm_ut[50 + m];
  const int f2 = sm_ut[75 + m];
  const int f4 = sm_ut[100 + m];
  in.pc4 = P->u_pcs[3LL * 650LL + i];
  const int f7 = sm_ut[125 + m];
  in.pc7 = P->u_pcs[4LL * 650LL + i];
  const int f10 = sm_ut[150 + m];
  const int f13 = sm_ut[175 + m];
  const int f16 = sm_ut[200 + m];
  in.pc16 = P->u_pcs[7LL * 650LL + i];
  in.x6 = P->u_pay1[0LL * 650LL + i];
  in.x9 = P->u_pay1[1LL * 650LL + i];
  in.x12 = P->u_pay1[2LL * 650LL + i];
  in.x15 = P->u_pay1[3LL * 650LL + i];
  in.x18 = P->u_pay1[4LL * 650LL + i];
  in.x21 = P->u_pay1[5LL * 650LL + i];
  in.s1 = __ldcg(sig_in + 0LL * 47008LL + f0);
  in.s2 = __ldcg(sig_in + 1LL * 47008LL + f0);
  in.s3 = __ldcg(sig_in + 0LL * 47008LL + f1);
  in.s4 = __ldcg(sig_in + 1LL * 47008LL + f1);
  in.s5 = __ldcg(sig_in + 0LL * 47008LL + f2);
  in.s6 = __ldcg(sig_in + 1LL * 47008LL + f2);
  in.s7 = __ldcg(sig_in + 2LL * 47008LL + f2);
  in.s8 = __ldcg(sig_in + 0LL * 47008LL + f4);
  in.s9 = __ldcg(sig_in + 1LL * 47008LL + f4);
  in.s10 = __ldcg(sig_in + 2LL * 47008LL + f4);
  in.s11 = __ldcg(sig_in + 0LL * 47008LL + f7);
  in.s12 = __ldcg(sig_in + 1LL * 47008LL + f7);
  in.s13 = __ldcg(sig_in + 2LL * 47008LL + f7);
  in.s14 = __ldcg(sig_in + 0LL * 47008LL + f10);
  in.s15 = __ldcg(sig_in + 1LL * 47008LL + f10);
  in.s16 = __ldcg(sig_in + 2LL * 47008LL + f10);
  in.s17 = __ldcg(sig_in + 0LL * 47008LL + f13);
  in.s18 = __ldcg(sig_in + 1LL * 47008LL + f13);
  in.s19 = __ldcg(sig_in + 2LL * 47008LL + f13);
  in.s20 = __ldcg(sig_in + 0LL * 47008LL + f16);
  in.s21 = __ldcg(sig_in + 1LL * 47008LL + f16);
  in.s22 = __ldcg(sig_in + 2LL * 47008LL + f16);
  in.pad_ = 0.0;
}
__device__ __forceinline__ void upper_eval0(int m, double w, const UIn0& in, const double* sm_fan, double* sm_r, double* sm_o) {
  const double pc0 = in.pc0;
  const double s1 = in.s1;
  const double s2 = in.s2;
  const double s3 = in.s3;
  const double x3 = sm_fan[m * 9 + 0];
  const double pc4 = in.pc4;
  const double s4 = in.s4;
  const double s5 = in.s5;
  const double x5 = sm_fan[m * 9 + 1];
  const double x6 = in.x6;
  const double s6 = in.s6;
  const double pc7 = in.pc7;
  const double s7 = in.s7;
  const double s8 = in.s8;
  const double x8 = sm_fan[m * 9 + 2];
  const double x9 = in.x9;
  const double s9 = in.s9;
  const double s10 = in.s10;
  const double s11 = in.s11;
  const double x11 = sm_fan[m * 9 + 3];
  const double x12 = in.x12;
  const double s12 = in.s12;
  const double s13 = in.s13;
  const double s14 = in.s14;
  const double x14 = sm_fan[m * 9 + 4];
  const double x15 = in.x15;
  const double s15 = in.s15;
  const double pc16 = in.pc16;
  const double s16 = in.s16;
  const double s17 = in.s17;
  const double x17 = sm_fan[m * 9 + 5];
  const double x18 = in.x18;
  const double s18 = in.s18;
  const double s19 = in.s19;
  const double x19 = sm_fan[m * 9 + 6];
  const double s20 = in.s20;
  const double x20 = sm_fan[m * 9 + 7];
  const double x21 = in.x21;
  const double s21 = in.s21;
  const double s22 = in.s22;
  const double x22 = sm_fan[m * 9 + 8];
  const double q1_0 = 1.0, q2_0 = 1.0;
  const double q1_1 = s1 * q1_0;
  const double q2_1 = q2_0;
  const double q1_2 = s2 * q1_0;
  const double q2_2 = q2_0;
  const double q1_4 = q1_1;
  const double q2_4 = s4 * q2_1;
  const double q1_7 = q1_2;
  const double q2_7 = s7 * q2_2;
  const double q1_10 = s10 * q1_4;
  const double q2_10 = q2_4;
  const double q1_13 = s13 * q1_7;
  const double q2_13 = q2_7;
  const double q1_16 = q1_10;
  const double q2_16 = s16 * q2_10;
  double a16 = 0.0;
  a16 = a16 + s20 * x20;
  a16 = a16 + s21 * x21;
  a16 = a16 + s22 * x22;
  const double x16 = a16;
  {
    const double opp = pc16 * q2_16;
    const double own = q1_16;
    const double vp = x16;
    sm_r[8 * 25 + m] = w * (x20 - vp) * opp;
    sm_r[9 * 25 + m] = w * (x21 - vp) * opp;
    sm_r[10 * 25 + m] = w * (x22 - vp) * opp;
    sm_o[3 * 25 + m] = w * own;
  }
  double a13 = 0.0;
  a13 = a13 + s17 * x17;
  a13 = a13 + s18 * x18;
  a13 = a13 + s19 * x19;
  const double x13 = a13;
  double a10 = 0.0;
  a10 = a10 + s14 * x14;
  a10 = a10 + s15 * x15;
  a10 = a10 + s16 * x16;
  const double x10 = a10;
  double a7 = 0.0;
  a7 = a7 + s11 * x11;
  a7 = a7 + s12 * x12;
  a7 = a7 + s13 * x13;
  const double x7 = a7;
  {
    const double opp = pc7 * q2_7;
    const double own = q1_7;
    const double vp = x7;
    sm_r[5 * 25 + m] = w * (x11 - vp) * opp;
    sm_r[6 * 25 + m] = w * (x12 - vp) * opp;
    sm_r[7 * 25 + m] = w * (x13 - vp) * opp;
    sm_o[2 * 25 + m] = w * own;
  }
  double a4 = 0.0;
  a4 = a4 + s8 * x8;
  a4 = a4 + s9 * x9;
  a4 = a4 + s10 * x10;
  const double x4 = a4;
  {
    const double opp = pc4 * q2_4;
    const double own = q1_4;
    const double vp = x4;
    sm_r[2 * 25 + m] = w * (x8 - vp) * opp;
    sm_r[3 * 25 + m] = w * (x9 - vp) * opp;
    sm_r[4 * 25 + m] = w * (x10 - vp) * opp;
    sm_o[1 * 25 + m] = w * own;
  }
  double a2 = 0.0;
  a2 = a2 + s5 * x5;
  a2 = a2 + s6 * x6;
  a2 = a2 + s7 * x7;
  const double x2 = a2;
  double a1 = 0.0;
  a1 = a1 + s3 * x3;
  a1 = a1 + s4 * x4;
  const double x1 = a1;
  double a0 = 0.0;
  a0 = a0 + s1 * x1;
  a0 = a0 + s2 * x2;
  const double x0 = a0;
  {
    const double opp = pc0 * q2_0;
    const double own = q1_0;
    const double vp = x0;
    sm_r[0 * 25 + m] = w * (x1 - vp) * opp;
    sm_r[1 * 25 + m] = w * (x2 - vp) * opp;
    sm_o[0 * 25 + m] = w * own;
  }
  (void)x0; (void)q1_0; (void)q2_0;
}
__device__ __forceinline__ void upper_reach0(double* out, int u, const double* sm_s) {
  const double q_0 = 1.0;
  const double q_1 = sm_s[0] * q_0;
  const double q_2 = sm_s[1] * q_0;
  out[0 + u] = q_1;
  const double q_4 = q_1;
  out[624 + u] = q_2;
  const double q_7 = q_2;
  out[1248 + u] = sm_s[2] * q_4;
  const double q_10 = sm_s[4] * q_4;
  out[1872 + u] = sm_s[5] * q_7;
  const double q_13 = sm_s[7] * q_7;
  out[2496 + u] = q_10;
  const double q_16 = q_10;
  out[3120 + u] = q_13;
  out[3744 + u] = q_13;
  out[4368 + u] = sm_s[8] * q_16;
  out[4992 + u] = sm_s[10] * q_16;
}
__constant__ int MCJ1[11] = {0,0,1,1,1,2,2,2,3,3,3};
__constant__ int MCA1[11] = {0,1,0,1,2,0,1,2,0,1,2};
__constant__ int MCOFF1[4] = {0,2,5,8};
__constant__ int MNA1[4] = {2,3,3,3};
__constant__ int UCJ1[11] = {0,0,1,1,1,2,2,2,3,3,3};
__constant__ int UCA1[11] = {0,1,0,1,2,0,1,2,0,1,2};
__constant__ int UCOFF1[4] = {0,2,5,8};
__constant__ int UNA1[4] = {2,3,3,3};
__device__ __forceinline__ MIds micro_ids1(const FusedParams* __restrict__ P, long long g, int k, bool ok,
    long long ag, int aj, bool acc) {
  MIds id;
  const long long slot = ok ? g * 24 + k : 0;
  const long long gg = ok ? g : 0;
  id.pat = P->mem_pat1[slot]; id.ri = P->mem_ridx1[slot]; id.ro = P->grp_ridx1[gg];
  id.pcv = P->mem_pc1[slot];
  id.pi = P->mem_pidx1[slot];
  id.f[0] = P->grp_fid1[gg * 4 + 0];
  id.f[1] = P->mem_opp1[0LL * 140400LL + slot];
  id.f[2] = P->mem_opp1[1LL * 140400LL + slot];
  id.f[3] = P->grp_fid1[gg * 4 + 1];
  id.f[4] = P->grp_fid1[gg * 4 + 2];
  id.f[5] = P->mem_opp1[2LL * 140400LL + slot];
  id.f[6] = P->mem_opp1[3LL * 140400LL + slot];
  id.f[7] = P->grp_fid1[gg * 4 + 3];
  const long long agg = acc ? ag : 0;
  id.acol = P->grp_col1[agg * 4 + aj];
  id.afid = P->grp_fid1[agg * 4 + aj];
  return id;
}
__device__ __forceinline__ void micro_eval1(const FusedParams* __restrict__ P, const MIds& id, int tix, double w,
    const double* sig_in, double* sm_r, double* sm_o, int par) {
  const int pat = id.pat, ri = id.ri, ro = id.ro;
  const double pcv = id.pcv;
  const int pi = id.pi;
  const int f0 = id.f[0];
  const int f1 = id.f[1];
  const int f2 = id.f[2];
  const int f3 = id.f[3];
  const int f4 = id.f[4];
  const int f5 = id.f[5];
  const int f6 = id.f[6];
  const int f7 = id.f[7];
  const int4* pp = reinterpret_cast<const int4*>(P->paytab8 + (long long)pat * 16);
  const int4 pw0 = __ldg(pp + 0);
  const double x3 = (double)(int)(signed char)(pw0.x >> 0);
  const double x5 = (double)(int)(signed char)(pw0.x >> 8);
  const double x6 = (double)(int)(signed char)(pw0.x >> 16);
  const double x8 = (double)(int)(signed char)(pw0.x >> 24);
  const double x9 = (double)(int)(signed char)(pw0.y >> 0);
  const double x11 = (double)(int)(signed char)(pw0.y >> 8);
  const double x12 = (double)(int)(signed char)(pw0.y >> 16);
  const double x14 = (double)(int)(signed char)(pw0.y >> 24);
  const double x15 = (double)(int)(signed char)(pw0.z >> 0);
  const double x17 = (double)(int)(signed char)(pw0.z >> 8);
  const double x18 = (double)(int)(signed char)(pw0.z >> 16);
  const double x19 = (double)(int)(signed char)(pw0.z >> 24);
  const double x20 = (double)(int)(signed char)(pw0.w >> 0);
  const double x21 = (double)(int)(signed char)(pw0.w >> 8);
  const double x22 = (double)(int)(signed char)(pw0.w >> 16);
  const double q1_0 = __ldcg(P->portal_p1 + ri), q2_0 = __ldcg(P->portal_p2 + ro);
  const double s1 = __ldcg(sig_in + 0LL * 47008LL + f0);
  const double s2 = __ldcg(sig_in + 1LL * 47008LL + f0);
  const double s3 = __ldcg(sig_in + 0LL * 47008LL + f1);
  const double s4 = __ldcg(sig_in + 1LL * 47008LL + f1);
  const double s5 = __ldcg(sig_in + 0LL * 47008LL + f2);
  const double s6 = __ldcg(sig_in + 1LL * 47008LL + f2);
  const double s7 = __ldcg(sig_in + 2LL * 47008LL + f2);
  const double s8 = __ldcg(sig_in + 0LL * 47008LL + f3);
  const double s9 = __ldcg(sig_in + 1LL * 47008LL + f3);
  const double s10 = __ldcg(sig_in + 2LL * 47008LL + f3);
  const double s11 = __ldcg(sig_in + 0LL * 47008LL + f4);
  const double s12 = __ldcg(sig_in + 1LL * 47008LL + f4);
  const double s13 = __ldcg(sig_in + 2LL * 47008LL + f4);
  const double s14 = __ldcg(sig_in + 0LL * 47008LL + f5);
  const double s15 = __ldcg(sig_in + 1LL * 47008LL + f5);
  const double s16 = __ldcg(sig_in + 2LL * 47008LL + f5);
  const double s17 = __ldcg(sig_in + 0LL * 47008LL + f6);
  const double s18 = __ldcg(sig_in + 1LL * 47008LL + f6);
  const double s19 = __ldcg(sig_in + 2LL * 47008LL + f6);
  const double s20 = __ldcg(sig_in + 0LL * 47008LL + f7);
  const double s21 = __ldcg(sig_in + 1LL * 47008LL + f7);
  const double s22 = __ldcg(sig_in + 2LL * 47008LL + f7);
  const double q1_1 = q1_0;
  const double q2_1 = s1 * q2_0;
  const double q1_2 = q1_0;
  const double q2_2 = s2 * q2_0;
  const double q1_3 = s4 * q1_1;
  const double q2_3 = q2_1;
  const double q1_4 = s7 * q1_2;
  const double q2_4 = q2_2;
  const double q1_5 = q1_3;
  const double q2_5 = s10 * q2_3;
  const double q1_6 = q1_4;
  const double q2_6 = s13 * q2_4;
  const double q1_7 = s16 * q1_5;
  const double q2_7 = q2_5;
  double a16 = 0.0;
  a16 = a16 + s20 * x20;
  a16 = a16 + s21 * x21;
  a16 = a16 + s22 * x22;
  const double x16 = a16;
  {
    const double opp = pcv * q1_7;
    const double own = q2_7;
    const double vp = -x16;
    sm_r[8 * SMS + tix] = w * (-x20 - vp) * opp;
    sm_r[9 * SMS + tix] = w * (-x21 - vp) * opp;
    sm_r[10 * SMS + tix] = w * (-x22 - vp) * opp;
    sm_o[3 * SMS + tix] = w * own;
  }
  double a13 = 0.0;
  a13 = a13 + s17 * x17;
  a13 = a13 + s18 * x18;
  a13 = a13 + s19 * x19;
  const double x13 = a13;
  double a10 = 0.0;
  a10 = a10 + s14 * x14;
  a10 = a10 + s15 * x15;
  a10 = a10 + s16 * x16;
  const double x10 = a10;
  double a7 = 0.0;
  a7 = a7 + s11 * x11;
  a7 = a7 + s12 * x12;
  a7 = a7 + s13 * x13;
  const double x7 = a7;
  {
    const double opp = pcv * q1_4;
    const double own = q2_4;
    const double vp = -x7;
    sm_r[5 * SMS + tix] = w * (-x11 - vp) * opp;
    sm_r[6 * SMS + tix] = w * (-x12 - vp) * opp;
    sm_r[7 * SMS + tix] = w * (-x13 - vp) * opp;
    sm_o[2 * SMS + tix] = w * own;
  }
  double a4 = 0.0;
  a4 = a4 + s8 * x8;
  a4 = a4 + s9 * x9;
  a4 = a4 + s10 * x10;
  const double x4 = a4;
  {
    const double opp = pcv * q1_3;
    const double own = q2_3;
    const double vp = -x4;
    sm_r[2 * SMS + tix] = w * (-x8 - vp) * opp;
    sm_r[3 * SMS + tix] = w * (-x9 - vp) * opp;
    sm_r[4 * SMS + tix] = w * (-x10 - vp) * opp;
    sm_o[1 * SMS + tix] = w * own;
  }
  double a2 = 0.0;
  a2 = a2 + s5 * x5;
  a2 = a2 + s6 * x6;
  a2 = a2 + s7 * x7;
  const double x2 = a2;
  double a1 = 0.0;
  a1 = a1 + s3 * x3;
  a1 = a1 + s4 * x4;
  const double x1 = a1;
  double a0 = 0.0;
  a0 = a0 + s1 * x1;
  a0 = a0 + s2 * x2;
  const double x0 = a0;
  {
    const double opp = pcv * q1_0;
    const double own = q2_0;
    const double vp = -x0;
    sm_r[0 * SMS + tix] = w * (-x1 - vp) * opp;
    sm_r[1 * SMS + tix] = w * (-x2 - vp) * opp;
    sm_o[0 * SMS + tix] = w * own;
  }
  { double* const* pvp = par ? &P->pvB0 : &P->pvA0; const int nr = (int)P->nranks;
    for (int r = 0; r < nr; ++r) pvp[r][pi] = x0; }
}
struct UIn1 {
  double pc1;
  double s1;
  double pc2;
  double s2;
  double s3;
  double s4;
  double s5;
  double x6;
  double s6;
  double s7;
  double s8;
  double x9;
  double s9;
  double pc10;
  double s10;
  double s11;
  double x12;
  double s12;
  double pc13;
  double s13;
  double s14;
  double x15;
  double s15;
  double s16;
  double s17;
  double x18;
  double s18;
  double s19;
  double s20;
  double x21;
  double s21;
  double s22;
  double pad_;
};
__device__ __forceinline__ void upper_load1(const FusedParams* __restrict__ P, long long i, int m, const int* sm_ut, const double* sig_in, const double* blk, UIn1& in) {
  const int f0 = sm_ut[25 + m];
  const int f1 = sm_ut[50 + m];
  in.pc1 = P->u_pcs[1LL * 650LL + i];
  const int f2 = sm_ut[75 + m];
  in.pc2 = P->u_pcs[2LL * 650LL + i];
  const int f4 = sm_ut[100 + m];
  const int f7 = sm_ut[125 + m];
  const int f10 = sm_ut[150 + m];
  in.pc10 = P->u_pcs[5LL * 650LL + i];
  const int f13 = sm_ut[175 + m];
  in.pc13 = P->u_pcs[6LL * 650LL + i];
  const int f16 = sm_ut[200 + m];
  in.x6 = P->u_pay1[0LL * 650LL + i];
  in.x9 = P->u_pay1[1LL * 650LL + i];
  in.x12 = P->u_pay1[2LL * 650LL + i];
  in.x15 = P->u_pay1[3LL * 650LL + i];
  in.x18 = P->u_pay1[4LL * 650LL + i];
  in.x21 = P->u_pay1[5LL * 650LL + i];
  in.s1 = __ldcg(sig_in + 0LL * 47008LL + f0);
  in.s2 = __ldcg(sig_in + 1LL * 47008LL + f0);
  in.s3 = __ldcg(sig_in + 0LL * 47008LL + f1);
  in.s4 = __ldcg(sig_in + 1LL * 47008LL + f1);
  in.s5 = __ldcg(sig_in + 0LL * 47008LL + f2);
  in.s6 = __ldcg(sig_in + 1LL * 47008LL + f2);
  in.s7 = __ldcg(sig_in + 2LL * 47008LL + f2);
  in.s8 = __ldcg(sig_in + 0LL * 47008LL + f4);
  in.s9 = __ldcg(sig_in + 1LL * 47008LL + f4);
  in.s10 = __ldcg(sig_in + 2LL * 47008LL + f4);
  in.s11 = __ldcg(sig_in + 0LL * 47008LL + f7);
  in.s12 = __ldcg(sig_in + 1LL * 47008LL + f7);
  in.s13 = __ldcg(sig_in + 2LL * 47008LL + f7);
  in.s14 = __ldcg(sig_in + 0LL * 47008LL + f10);
  in.s15 = __ldcg(sig_in + 1LL * 47008LL + f10);
  in.s16 = __ldcg(sig_in + 2LL * 47008LL + f10);
  in.s17 = __ldcg(sig_in + 0LL * 47008LL + f13);
  in.s18 = __ldcg(sig_in + 1LL * 47008LL + f13);
  in.s19 = __ldcg(sig_in + 2LL * 47008LL + f13);
  in.s20 = __ldcg(sig_in + 0LL * 47008LL + f16);
  in.s21 = __ldcg(sig_in + 1LL * 47008LL + f16);
  in.s22 = __ldcg(sig_in + 2LL * 47008LL + f16);
  in.pad_ = 0.0;
}
__device__ __forceinline__ void upper_eval1(int m, double w, const UIn1& in, const double* sm_fan, double* sm_r, double* sm_o) {
  const double pc1 = in.pc1;
  const double s1 = in.s1;
  const double pc2 = in.pc2;
  const double s2 = in.s2;
  const double s3 = in.s3;
  const double x3 = sm_fan[m * 9 + 0];
  const double s4 = in.s4;
  const double s5 = in.s5;
  const double x5 = sm_fan[m * 9 + 1];
  const double x6 = in.x6;
  const double s6 = in.s6;
  const double s7 = in.s7;
  const double s8 = in.s8;
  const double x8 = sm_fan[m * 9 + 2];
  const double x9 = in.x9;
  const double s9 = in.s9;
  const double pc10 = in.pc10;
  const double s10 = in.s10;
  const double s11 = in.s11;
  const double x11 = sm_fan[m * 9 + 3];
  const double x12 = in.x12;
  const double s12 = in.s12;
  const double pc13 = in.pc13;
  const double s13 = in.s13;
  const double s14 = in.s14;
  const double x14 = sm_fan[m * 9 + 4];
  const double x15 = in.x15;
  const double s15 = in.s15;
  const double s16 = in.s16;
  const double s17 = in.s17;
  const double x17 = sm_fan[m * 9 + 5];
  const double x18 = in.x18;
  const double s18 = in.s18;
  const double s19 = in.s19;
  const double x19 = sm_fan[m * 9 + 6];
  const double s20 = in.s20;
  const double x20 = sm_fan[m * 9 + 7];
  const double x21 = in.x21;
  const double s21 = in.s21;
  const double s22 = in.s22;
  const double x22 = sm_fan[m * 9 + 8];
  const double q1_0 = 1.0, q2_0 = 1.0;
  const double q1_1 = s1 * q1_0;
  const double q2_1 = q2_0;
  const double q1_2 = s2 * q1_0;
  const double q2_2 = q2_0;
  const double q1_4 = q1_1;
  const double q2_4 = s4 * q2_1;
  const double q1_7 = q1_2;
  const double q2_7 = s7 * q2_2;
  const double q1_10 = s10 * q1_4;
  const double q2_10 = q2_4;
  const double q1_13 = s13 * q1_7;
  const double q2_13 = q2_7;
  const double q1_16 = q1_10;
  const double q2_16 = s16 * q2_10;
  double a16 = 0.0;
  a16 = a16 + s20 * x20;
  a16 = a16 + s21 * x21;
  a16 = a16 + s22 * x22;
  const double x16 = a16;
  double a13 = 0.0;
  a13 = a13 + s17 * x17;
  a13 = a13 + s18 * x18;
  a13 = a13 + s19 * x19;
  const double x13 = a13;
  {
    const double opp = pc13 * q1_13;
    const double own = q2_13;
    const double vp = -x13;
    sm_r[8 * 25 + m] = w * (-x17 - vp) * opp;
    sm_r[9 * 25 + m] = w * (-x18 - vp) * opp;
    sm_r[10 * 25 + m] = w * (-x19 - vp) * opp;
    sm_o[3 * 25 + m] = w * own;
  }
  double a10 = 0.0;
  a10 = a10 + s14 * x14;
  a10 = a10 + s15 * x15;
  a10 = a10 + s16 * x16;
  const double x10 = a10;
  {
    const double opp = pc10 * q1_10;
    const double own = q2_10;
    const double vp = -x10;
    sm_r[5 * 25 + m] = w * (-x14 - vp) * opp;
    sm_r[6 * 25 + m] = w * (-x15 - vp) * opp;
    sm_r[7 * 25 + m] = w * (-x16 - vp) * opp;
    sm_o[2 * 25 + m] = w * own;
  }
  double a7 = 0.0;
  a7 = a7 + s11 * x11;
  a7 = a7 + s12 * x12;
  a7 = a7 + s13 * x13;
  const double x7 = a7;
  double a4 = 0.0;
  a4 = a4 + s8 * x8;
  a4 = a4 + s9 * x9;
  a4 = a4 + s10 * x10;
  const double x4 = a4;
  double a2 = 0.0;
  a2 = a2 + s5 * x5;
  a2 = a2 + s6 * x6;
  a2 = a2 + s7 * x7;
  const double x2 = a2;
  {
    const double opp = pc2 * q1_2;
    const double own = q2_2;
    const double vp = -x2;
    sm_r[2 * 25 + m] = w * (-x5 - vp) * opp;
    sm_r[3 * 25 + m] = w * (-x6 - vp) * opp;
    sm_r[4 * 25 + m] = w * (-x7 - vp) * opp;
    sm_o[1 * 25 + m] = w * own;
  }
  double a1 = 0.0;
  a1 = a1 + s3 * x3;
  a1 = a1 + s4 * x4;
  const double x1 = a1;
  {
    const double opp = pc1 * q1_1;
    const double own = q2_1;
    const double vp = -x1;
    sm_r[0 * 25 + m] = w * (-x3 - vp) * opp;
    sm_r[1 * 25 + m] = w * (-x4 - vp) * opp;
    sm_o[0 * 25 + m] = w * own;
  }
  double a0 = 0.0;
  a0 = a0 + s1 * x1;
  a0 = a0 + s2 * x2;
  const double x0 = a0;
  (void)x0; (void)q1_0; (void)q2_0;
}
__device__ __forceinline__ void upper_reach1(double* out, int u, const double* sm_s) {
  const double q_0 = 1.0;
  const double q_1 = q_0;
  const double q_2 = q_0;
  out[0 + u] = sm_s[0] * q_1;
  const double q_4 = sm_s[1] * q_1;
  out[624 + u] = sm_s[2] * q_2;
  const double q_7 = sm_s[4] * q_2;
  out[1248 + u] = q_4;
  const double q_10 = q_4;
  out[1872 + u] = q_7;
  const double q_13 = q_7;
  out[2496 + u] = sm_s[5] * q_10;
  const double q_16 = sm_s[7] * q_10;
  out[3120 + u] = sm_s[8] * q_13;
  out[3744 + u] = sm_s[10] * q_13;
  out[4368 + u] = q_16;
  out[4992 + u] = q_16;
}
__device__ __forceinline__ MIds micro_unit_ids0(const FusedParams* __restrict__ P, long long unit) {
  const int tix = threadIdx.x % TEAM;
  const int gl = tix / 24, k = tix - gl * 24;
  const long long g0 = P->first_group0 + unit * 10, gend = P->end_group0;
  const long long g = g0 + gl;
  const int agl = tix / 11, ci = tix - agl * 11;
  const bool acc = tix < 110 && g0 + agl < gend;
  return micro_ids0(P, g, k, gl < 10 && g < gend, g0 + agl, acc ? MCJ0[ci] : 0, acc);
}
__device__ __forceinline__ void micro_unit0(const FusedParams* __restrict__ P, long long unit, const MIds& id, double w, const double* sig_in,
    double* sig_out, double* sm_r, double* sm_o, double* sm_rn, int par) {
  const int tix = threadIdx.x % TEAM;
  const int gl = tix / 24;
  const long long g0 = P->first_group0 + unit * 10, gend = P->end_group0;
  const long long g = g0 + gl;
  const int agl = tix / 11, ci = tix - agl * 11;
  const long long ag = g0 + agl;
  const bool acc = tix < 110 && ag < gend;
  int j = 0, a = 0, col = 0, fid = 0;
  double r = 0.0, sa = 0.0, sg = 0.0;
  if (acc) {
    j = MCJ0[ci]; a = MCA0[ci];
    col = id.acol + a;
    fid = id.afid;
    r = __ldcg(P->R + col); sa = __ldcg(P->S + col);
    sg = __ldcg(sig_in + (long long)a * 47008LL + fid);
  }
  if (gl < 10 && g < gend) micro_eval0(P, id, tix, w, sig_in, sm_r, sm_o, par);
  TEAM_SYNC();
  if (acc) {
    const double* cr = sm_r + ci * SMS + agl * 24;
    const double* co = sm_o + j * SMS + agl * 24;
#pragma unroll 8
    for (int m = 0; m < 24; ++m) { r = r + cr[m]; sa = sa + co[m] * sg; }
    P->R[col] = r; P->S[col] = sa;
    sm_rn[tix] = r;
  }
  TEAM_SYNC();
  if (acc) {
    const double out = rm_one(sm_rn + agl * 11 + MCOFF0[j], MNA0[j], a);
    P->sigma_row[col] = out;
    sig_out[(long long)a * 47008LL + fid] = out;
    if (a == 0) { const int is = P->grp_iset0[ag * 4 + j]; const long long I = P->num_isets;
      P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }
  }
  TEAM_SYNC();
}
__device__ __noinline__ void upper_cache0(const FusedParams* __restrict__ P, int u, int* sm_ut) {
  const int tid = threadIdx.x;
  for (int x = tid; x < 25; x += BLOCK) sm_ut[x] = P->tg_tile0[(long long)u * 25 + x];
  for (int x = tid; x < 200; x += BLOCK) {
    const int d = x / 25, m = x - d * 25;
    sm_ut[25 + d * 25 + m] = P->u_fid[(long long)d * 650LL + P->tg_tile0[(long long)u * 25 + m]];
  }
  for (int x = tid; x < 4; x += BLOCK) {
    sm_ut[225 + x] = P->tg_col0[u * 4 + x];
    sm_ut[229 + x] = P->tg_fid0[u * 4 + x];
    sm_ut[233 + x] = P->tg_iset0[u * 4 + x];
  }
}
__device__ __noinline__ void upper_unit0(const FusedParams* __restrict__ P, int u, double w, const double* sig_in,
    double* sig_out, int reach_only, double* sm_fan, double* sm_r, double* sm_o, double* sm_rn, double* sm_s, const int* sm_ut,
    int par, unsigned long long* st) {
  const int tid = threadIdx.x;
  const int* tiles = sm_ut;
  if (reach_only) {
    for (int x = tid; x < 11; x += BLOCK) {
      const int j = UCJ0[x], a = UCA0[x];
      sm_s[x] = __ldcg(sig_in + (long long)a * 47008LL + sm_ut[229 + j]);
    }
  } else {
    int aj = 0, aa = 0, acol = 0, afid = 0;
    double ar = 0.0, asa = 0.0, asg = 0.0;
    if (tid < 11) {
      aj = UCJ0[tid]; aa = UCA0[tid];
      acol = sm_ut[225 + aj] + aa;
      afid = sm_ut[229 + aj];
      ar = __ldcg(P->R + acol); asa = __ldcg(P->S + acol);
      asg = __ldcg(sig_in + (long long)aa * 47008LL + afid);
    }
    const int* fd = P->fan_desc;
    const double* blk = (par ? P->pvl1 : P->pvl0) + 0LL + (long long)u * 5400LL;
    UIn0 uin;
    if (tid < 25) upper_load0(P, tiles[tid], tid, sm_ut, sig_in, blk, uin);
    if (st && tid == 0) st[0] = now_ns();
    double fv[22];
#pragma unroll
    for (int i = 0; i < 22; ++i) { const int L = tid + i * BLOCK; if (L < 5400) fv[i] = __ldcg(blk + L); }
    {
#pragma unroll
      for (int i = 0; i < 22; ++i) { const int L = tid + i * BLOCK; if (L >= 0 && L < 3600) sm_r[L - 0] = fv[i]; }
      __syncthreads();
      for (int x = tid; x < 150; x += BLOCK) {
        const int fl = x / 25, m = x - fl * 25, f = 0 + fl;
        const int k0 = fd[4 * f], nch = fd[4 * f + 1];
        const double e = FCP[f];
        const double* pv = sm_r + (k0 - 0) * 25 + m;
        double acc = 0.0;
        for (int ch = 0; ch < nch; ++ch) acc = acc + e * pv[ch * 25];
        sm_fan[m * 9 + f] = acc;
      }
      __syncthreads();
    }
    {
#pragma unroll
      for (int i = 0; i < 22; ++i) { const int L = tid + i * BLOCK; if (L >= 3600 && L < 5400) sm_r[L - 3600] = fv[i]; }
      __syncthreads();
      for (int x = tid; x < 75; x += BLOCK) {
        const int fl = x / 25, m = x - fl * 25, f = 6 + fl;
        const int k0 = fd[4 * f], nch = fd[4 * f + 1];
        const double e = FCP[f];
        const double* pv = sm_r + (k0 - 144) * 25 + m;
        double acc = 0.0;
        for (int ch = 0; ch < nch; ++ch) acc = acc + e * pv[ch * 25];
        sm_fan[m * 9 + f] = acc;
      }
      __syncthreads();
    }
    if (st && tid == 0) st[1] = now_ns();
    if (tid < 25) upper_eval0(tid, w, uin, sm_fan, sm_r, sm_o);
    __syncthreads();
    if (st && tid == 0) st[2] = now_ns();
    if (tid < 11) {
      const double* cr = sm_r + tid * 25;
      const double* co = sm_o + aj * 25;
      for (int m = 0; m < 25; ++m) { ar = ar + cr[m]; asa = asa + co[m] * asg; }
      P->R[acol] = ar; P->S[acol] = asa;
      sm_rn[tid] = ar;
    }
    __syncthreads();
    if (tid < 11) {
      const double out = rm_one(sm_rn + UCOFF0[aj], UNA0[aj], aa);
      P->sigma_row[acol] = out;
      sig_out[(long long)aa * 47008LL + afid] = out;
      sm_s[tid] = out;
      if (aa == 0) { const int is = sm_ut[233 + aj]; const long long I = P->num_isets;
        P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }
    }
  }
  __syncthreads();
  if (st && tid == 0) st[3] = now_ns();
  if (tid == 0) upper_reach0(P->portal_p1, u, sm_s);
  __syncthreads();
  if (st && tid == 0) st[4] = now_ns();
}
__device__ __forceinline__ MIds micro_unit_ids1(const FusedParams* __restrict__ P, long long unit) {
  const int tix = threadIdx.x % TEAM;
  const int gl = tix / 24, k = tix - gl * 24;
  const long long g0 = P->first_group1 + unit * 10, gend = P->end_group1;
  const long long g = g0 + gl;
  const int agl = tix / 11, ci = tix - agl * 11;
  const bool acc = tix < 110 && g0 + agl < gend;
  return micro_ids1(P, g, k, gl < 10 && g < gend, g0 + agl, acc ? MCJ1[ci] : 0, acc);
}
__device__ __forceinline__ void micro_unit1(const FusedParams* __restrict__ P, long long unit, const MIds& id, double w, const double* sig_in,
    double* sig_out, double* sm_r, double* sm_o, double* sm_rn, int par) {
  const int tix = threadIdx.x % TEAM;
  const int gl = tix / 24;
  const long long g0 = P->first_group1 + unit * 10, gend = P->end_group1;
  const long long g = g0 + gl;
  const int agl = tix / 11, ci = tix - agl * 11;
  const long long ag = g0 + agl;
  const bool acc = tix < 110 && ag < gend;
  int j = 0, a = 0, col = 0, fid = 0;
  double r = 0.0, sa = 0.0, sg = 0.0;
  if (acc) {
    j = MCJ1[ci]; a = MCA1[ci];
    col = id.acol + a;
    fid = id.afid;
    r = __ldcg(P->R + col); sa = __ldcg(P->S + col);
    sg = __ldcg(sig_in + (long long)a * 47008LL + fid);
  }
  if (gl < 10 && g < gend) micro_eval1(P, id, tix, w, sig_in, sm_r, sm_o, par);
  TEAM_SYNC();
  if (acc) {
    const double* cr = sm_r + ci * SMS + agl * 24;
    const double* co = sm_o + j * SMS + agl * 24;
#pragma unroll 8
    for (int m = 0; m < 24; ++m) { r = r + cr[m]; sa = sa + co[m] * sg; }
    P->R[col] = r; P->S[col] = sa;
    sm_rn[tix] = r;
  }
  TEAM_SYNC();
  if (acc) {
    const double out = rm_one(sm_rn + agl * 11 + MCOFF1[j], MNA1[j], a);
    P->sigma_row[col] = out;
    sig_out[(long long)a * 47008LL + fid] = out;
    if (a == 0) { const int is = P->grp_iset1[ag * 4 + j]; const long long I = P->num_isets;
      P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }
  }
  TEAM_SYNC();
}
__device__ __noinline__ void upper_cache1(const FusedParams* __restrict__ P, int u, int* sm_ut) {
  const int tid = threadIdx.x;
  for (int x = tid; x < 25; x += BLOCK) sm_ut[x] = P->tg_tile1[(long long)u * 25 + x];
  for (int x = tid; x < 200; x += BLOCK) {
    const int d = x / 25, m = x - d * 25;
    sm_ut[25 + d * 25 + m] = P->u_fid[(long long)d * 650LL + P->tg_tile1[(long long)u * 25 + m]];
  }
  for (int x = tid; x < 4; x += BLOCK) {
    sm_ut[225 + x] = P->tg_col1[u * 4 + x];
    sm_ut[229 + x] = P->tg_fid1[u * 4 + x];
    sm_ut[233 + x] = P->tg_iset1[u * 4 + x];
  }
}
__device__ __noinline__ void upper_unit1(const FusedParams* __restrict__ P, int u, double w, const double* sig_in,
    double* sig_out, int reach_only, double* sm_fan, double* sm_r, double* sm_o, double* sm_rn, double* sm_s, const int* sm_ut,
    int par, unsigned long long* st) {
  const int tid = threadIdx.x;
  const int* tiles = sm_ut;
  if (reach_only) {
    for (int x = tid; x < 11; x += BLOCK) {
      const int j = UCJ1[x], a = UCA1[x];
      sm_s[x] = __ldcg(sig_in + (long long)a * 47008LL + sm_ut[229 + j]);
    }
  } else {
    int aj = 0, aa = 0, acol = 0, afid = 0;
    double ar = 0.0, asa = 0.0, asg = 0.0;
    if (tid < 11) {
      aj = UCJ1[tid]; aa = UCA1[tid];
      acol = sm_ut[225 + aj] + aa;
      afid = sm_ut[229 + aj];
      ar = __ldcg(P->R + acol); asa = __ldcg(P->S + acol);
      asg = __ldcg(sig_in + (long long)aa * 47008LL + afid);
    }
    const int* fd = P->fan_desc;
    const double* blk = (par ? P->pvl1 : P->pvl0) + 140400LL + (long long)u * 5400LL;
    UIn1 uin;
    if (tid < 25) upper_load1(P, tiles[tid], tid, sm_ut, sig_in, blk, uin);
    if (st && tid == 0) st[0] = now_ns();
    double fv[22];
#pragma unroll
    for (int i = 0; i < 22; ++i) { const int L = tid + i * BLOCK; if (L < 5400) fv[i] = __ldcg(blk + L); }
    {
#pragma unroll
      for (int i = 0; i < 22; ++i) { const int L = tid + i * BLOCK; if (L >= 0 && L < 3600) sm_r[L - 0] = fv[i]; }
      __syncthreads();
      for (int x = tid; x < 150; x += BLOCK) {
        const int fl = x / 25, m = x - fl * 25, f = 0 + fl;
        const int k0 = fd[4 * f], nch = fd[4 * f + 1];
        const double e = FCP[f];
        const double* pv = sm_r + (k0 - 0) * 25 + m;
        double acc = 0.0;
        for (int ch = 0; ch < nch; ++ch) acc = acc + e * pv[ch * 25];
        sm_fan[m * 9 + f] = acc;
      }
      __syncthreads();
    }
    {
#pragma unroll
      for (int i = 0; i < 22; ++i) { const int L = tid + i * BLOCK; if (L >= 3600 && L < 5400) sm_r[L - 3600] = fv[i]; }
      __syncthreads();
      for (int x = tid; x < 75; x += BLOCK) {
        const int fl = x / 25, m = x - fl * 25, f = 6 + fl;
        const int k0 = fd[4 * f], nch = fd[4 * f + 1];
        const double e = FCP[f];
        const double* pv = sm_r + (k0 - 144) * 25 + m;
        double acc = 0.0;
        for (int ch = 0; ch < nch; ++ch) acc = acc + e * pv[ch * 25];
        sm_fan[m * 9 + f] = acc;
      }
      __syncthreads();
    }
    if (st && tid == 0) st[1] = now_ns();
    if (tid < 25) upper_eval1(tid, w, uin, sm_fan, sm_r, sm_o);
    __syncthreads();
    if (st && tid == 0) st[2] = now_ns();
    if (tid < 11) {
      const double* cr = sm_r + tid * 25;
      const double* co = sm_o + aj * 25;
      for (int m = 0; m < 25; ++m) { ar = ar + cr[m]; asa = asa + co[m] * asg; }
      P->R[acol] = ar; P->S[acol] = asa;
      sm_rn[tid] = ar;
    }
    __syncthreads();
    if (tid < 11) {
      const double out = rm_one(sm_rn + UCOFF1[aj], UNA1[aj], aa);
      P->sigma_row[acol] = out;
      sig_out[(long long)aa * 47008LL + afid] = out;
      sm_s[tid] = out;
      if (aa == 0) { const int is = sm_ut[233 + aj]; const long long I = P->num_isets;
        P->flags[is] = 1; P->flags[I + is] = 1; P->flags[2 * I + is] = 1; }
    }
  }
  __syncthreads();
  if (st && tid == 0) st[3] = now_ns();
  if (tid == 0) upper_reach1(P->portal_p2, u, sm_s);
  __syncthreads();
  if (st && tid == 0) st[4] = now_ns();
}
extern "C" __global__ void __launch_bounds__(BLOCK, 3) cfr_fused_k(const __grid_constant__ FusedParams PV, long long t0, int iters,
    int linear, int cur0) {
  const FusedParams* __restrict__ P = &PV;      // parameter (constant) space: no global load behind a field read
  __shared__ double sm_ro[3855];
  __shared__ int sm_ut[237];
  double* sm_r = sm_ro;
  double* sm_o = sm_ro + 2827;
  int cached_u = -1;
  __shared__ double sm_rn[110];
  __shared__ double sm_fan[225];
  __shared__ double sm_s[11];
  __shared__ int s_lead;
  const unsigned int nb = gridDim.x;
  unsigned int nbar = 0;
  long long tlc = 0;
  const int NU0 = 26, NU1 = 26;
  // The upper phase has far fewer units than there are CTAs.  Consecutive CTAs share an SM, so the units go to
  // one LEADER CTA per SM (the first CTA to register on its SM), spread over all SMs.
  if (threadIdx.x == 0) {
    unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned int rank = atomicAdd(P->bar + 32 + 32 * nb + (smid & 1023u), 1u);
    s_lead = rank == 0 ? (int)atomicAdd(P->bar + 1, 1u) : -1;
  }
  __syncthreads();
  const int lead = s_lead;
  const bool logc = P->tl != 0 && lead == 0;            // the CTA whose view of the phases is recorded
  const bool log = logc && threadIdx.x == 0;
  // prologue: reach of both players to the portals from the current strategy
  {
    const double* sig_in = cur0 ? P->sigF1 : P->sigF0;
    for (int u = blockIdx.x; u < NU0 + NU1; u += nb) {
      if (u < NU0) upper_cache0(P, u, sm_ut); else upper_cache1(P, u - NU0, sm_ut);
      __syncthreads();
      if (u < NU0) upper_unit0(P, u, 1.0, sig_in, 0, 1, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, 0, 0);
      else upper_unit1(P, u - NU0, 1.0, sig_in, 0, 1, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, 0, 0);
    }
    grid_barrier(P->bar, ++nbar, nb);
  }
  const int n_lead = (int)__ldcg(P->bar + 1);
  const int team = threadIdx.x / TEAM;
  double* t_r = sm_r + team * 2827;
  double* t_o = sm_o + team * 1028;
  double* t_rn = sm_rn + team * 110;
  const long long nu0 = (P->end_group0 - P->first_group0 + 9) / 10;
  const long long nu1 = (P->end_group1 - P->first_group1 + 9) / 10;
  for (int it = 0; it < iters; ++it) {
    const int cur = (cur0 + it) & 1;
    const double* sig_in = cur ? P->sigF1 : P->sigF0;
    double* sig_out = cur ? P->sigF0 : P->sigF1;
    const double w = linear ? (double)(t0 + it) : 1.0;
    const int par = (int)((P->xbase + it) & 1);        // portal-value buffer of this iteration
    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();
    // phase A: micro units of player 1, then of player 2 (one list), one unit per team at a time
    {
      const long long stride = (long long)nb * 1;
      long long un = (long long)blockIdx.x * 1 + team;
      for (; un < nu0 + nu1; un += stride) {
        if (un < nu0) micro_unit0(P, un, micro_unit_ids0(P, un), w, sig_in, sig_out, t_r, t_o, t_rn, par);
        else micro_unit1(P, un - nu0, micro_unit_ids1(P, un - nu0), w, sig_in, sig_out, t_r, t_o, t_rn, par);
      }
    }
    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();
    // Multi-GPU: the micro phase has stored the values of this rank's subtrees into every rank's buffer; the barrier
    // that ends the phase also shakes hands with the peers (each CTA fences its peer stores at system scope first).
    if (P->nranks > 1) grid_barrier_x(P, ++nbar, nb, (unsigned long long)(P->xbase + it + 1), false);
    else grid_barrier(P->bar, ++nbar, nb);
    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();
    // phase B: upper groups (fan-in, tile, round-1 information sets, reach for the next iteration)
    if (lead >= 0) {
      for (int u = lead; u < NU0 + NU1; u += n_lead) {
        unsigned long long* st = (logc && u == lead && tlc + 5 <= P->tl_cap) ? P->tl + tlc : 0;
        if (u != cached_u) {          // a leader keeps its unit's static tables across iterations
          if (u < NU0) upper_cache0(P, u, sm_ut); else upper_cache1(P, u - NU0, sm_ut);
          cached_u = u;
          __syncthreads();
        }
        if (u < NU0) upper_unit0(P, u, w, sig_in, sig_out, 0, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, par, st);
        else upper_unit1(P, u - NU0, w, sig_in, sig_out, 0, sm_fan, sm_r, sm_o, sm_rn, sm_s, sm_ut, par, st);
      }
    }
    tlc += 5;
    if (log && tlc < P->tl_cap) P->tl[tlc++] = now_ns();
    grid_barrier(P->bar, ++nbar, nb);
  }
}
__device__ __forceinline__ void micro_br0(const FusedParams* __restrict__ P, long long unit, double* sm_q) {
  const int tix = threadIdx.x % TEAM;
  const int gl = tix / 24, k = tix - gl * 24;
  const long long g = unit * 10 + gl;
  const bool ok = gl < 10 && g < 5850LL;
  const int glr = gl < 10 ? gl : 0;          // column block read by the reductions (idle lanes stay in bounds)
  const long long slot = ok ? g * 24 + k : 0;
  const int pat = P->mem_pat0[slot], ri = P->mem_ridx0[slot], pi = P->mem_pidx0[slot];
  const double pcv = P->mem_pc0[slot];
  const int f0 = P->mem_opp0[0LL * 140400LL + slot];
  const int f3 = P->mem_opp0[1LL * 140400LL + slot];
  const int f4 = P->mem_opp0[2LL * 140400LL + slot];
  const int f7 = P->mem_opp0[3LL * 140400LL + slot];
  const int4* pp = reinterpret_cast<const int4*>(P->paytab8 + (long long)pat * 16);
  const int4 pw0 = __ldg(pp + 0);
  const double x3 = (double)(int)(signed char)(pw0.x >> 0);
  const double x5 = (double)(int)(signed char)(pw0.x >> 8);
  const double x6 = (double)(int)(signed char)(pw0.x >> 16);
  const double x8 = (double)(int)(signed char)(pw0.x >> 24);
  const double x9 = (double)(int)(signed char)(pw0.y >> 0);
  const double x11 = (double)(int)(signed char)(pw0.y >> 8);
  const double x12 = (double)(int)(signed char)(pw0.y >> 16);
  const double x14 = (double)(int)(signed char)(pw0.y >> 24);
  const double x15 = (double)(int)(signed char)(pw0.z >> 0);
  const double x17 = (double)(int)(signed char)(pw0.z >> 8);
  const double x18 = (double)(int)(signed char)(pw0.z >> 16);
  const double x19 = (double)(int)(signed char)(pw0.z >> 24);
  const double x20 = (double)(int)(signed char)(pw0.w >> 0);
  const double x21 = (double)(int)(signed char)(pw0.w >> 8);
  const double x22 = (double)(int)(signed char)(pw0.w >> 16);
  const double rho_0 = pcv * __ldcg(P->br_reach2 + ri);
  const double s1 = __ldcg(P->br_sig + 0LL * 47008LL + f0);
  const double s2 = __ldcg(P->br_sig + 1LL * 47008LL + f0);
  const double s8 = __ldcg(P->br_sig + 0LL * 47008LL + f3);
  const double s9 = __ldcg(P->br_sig + 1LL * 47008LL + f3);
  const double s10 = __ldcg(P->br_sig + 2LL * 47008LL + f3);
  const double s11 = __ldcg(P->br_sig + 0LL * 47008LL + f4);
  const double s12 = __ldcg(P->br_sig + 1LL * 47008LL + f4);
  const double s13 = __ldcg(P->br_sig + 2LL * 47008LL + f4);
  const double s20 = __ldcg(P->br_sig + 0LL * 47008LL + f7);
  const double s21 = __ldcg(P->br_sig + 1LL * 47008LL + f7);
  const double s22 = __ldcg(P->br_sig + 2LL * 47008LL + f7);
  const double rho_1 = rho_0 * s1;
  const double rho_2 = rho_0 * s2;
  const double rho_3 = rho_1;
  const double rho_4 = rho_2;
  const double rho_5 = rho_3 * s10;
  const double rho_6 = rho_4 * s13;
  const double rho_7 = rho_5;
  const double v20 = (rho_7 * s20) * x20;
  const double v21 = (rho_7 * s21) * x21;
  const double v22 = (rho_7 * s22) * x22;
  double V7 = 0.0;
  V7 = V7 + v20;
  V7 = V7 + v21;
  V7 = V7 + v22;
  const double v17 = (rho_6) * x17;
  const double v18 = (rho_6) * x18;
  const double v19 = (rho_6) * x19;
  if (ok) {
    sm_q[9 * SMS + tix] = v17;
    sm_q[10 * SMS + tix] = v18;
    sm_q[11 * SMS + tix] = v19;
  }
  TEAM_SYNC();
  double V6;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 9 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 10 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 11 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 2; } }
    V6 = v17;
    if (best == 1) V6 = v18;
    if (best == 2) V6 = v19;
    if (ok && k == 0) P->br_astar[P->grp_iset0[g * 4 + 3]] = best;
  }
  const double v14 = (rho_5) * x14;
  const double v15 = (rho_5) * x15;
  const double v16 = V7;
  if (ok) {
    sm_q[6 * SMS + tix] = v14;
    sm_q[7 * SMS + tix] = v15;
    sm_q[8 * SMS + tix] = v16;
  }
  TEAM_SYNC();
  double V5;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 6 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 7 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 8 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 2; } }
    V5 = v14;
    if (best == 1) V5 = v15;
    if (best == 2) V5 = v16;
    if (ok && k == 0) P->br_astar[P->grp_iset0[g * 4 + 2]] = best;
  }
  const double v11 = (rho_4 * s11) * x11;
  const double v12 = (rho_4 * s12) * x12;
  const double v13 = V6;
  double V4 = 0.0;
  V4 = V4 + v11;
  V4 = V4 + v12;
  V4 = V4 + v13;
  const double v8 = (rho_3 * s8) * x8;
  const double v9 = (rho_3 * s9) * x9;
  const double v10 = V5;
  double V3 = 0.0;
  V3 = V3 + v8;
  V3 = V3 + v9;
  V3 = V3 + v10;
  const double v5 = (rho_2) * x5;
  const double v6 = (rho_2) * x6;
  const double v7 = V4;
  if (ok) {
    sm_q[3 * SMS + tix] = v5;
    sm_q[4 * SMS + tix] = v6;
    sm_q[5 * SMS + tix] = v7;
  }
  TEAM_SYNC();
  double V2;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 3 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 4 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 5 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 2; } }
    V2 = v5;
    if (best == 1) V2 = v6;
    if (best == 2) V2 = v7;
    if (ok && k == 0) P->br_astar[P->grp_iset0[g * 4 + 1]] = best;
  }
  const double v3 = (rho_1) * x3;
  const double v4 = V3;
  if (ok) {
    sm_q[0 * SMS + tix] = v3;
    sm_q[1 * SMS + tix] = v4;
  }
  TEAM_SYNC();
  double V1;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 0 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 1 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    V1 = v3;
    if (best == 1) V1 = v4;
    if (ok && k == 0) P->br_astar[P->grp_iset0[g * 4 + 0]] = best;
  }
  const double v1 = V1;
  const double v2 = V2;
  double V0 = 0.0;
  V0 = V0 + v1;
  V0 = V0 + v2;
  if (ok) P->br_pv1[pi] = V0;
  TEAM_SYNC();
}
__device__ __noinline__ void upper_br0(const FusedParams* __restrict__ P, int u, double* sm_fan, double* sm_q, const int* sm_ut) {
  const int tid = threadIdx.x;
  const double* pvl = P->br_pv1 + 0LL + (long long)u * 5400LL;
  const int* fd = P->fan_desc;
  for (int x = tid; x < 225; x += BLOCK) {
    const int mm = x / 9, f = x - mm * 9;
    const int k0 = fd[4 * f], nch = fd[4 * f + 1];
    const double* pv = pvl + (long long)k0 * 25 + mm;
    double acc = 0.0;
    double vv[24];
#pragma unroll
    for (int ch = 0; ch < 24; ++ch) vv[ch] = __ldcg(pv + ch * 25);
#pragma unroll
    for (int ch = 0; ch < 24; ++ch) acc = acc + vv[ch];
    (void)nch;
    sm_fan[x] = acc;
  }
  __syncthreads();
  const bool act = tid < 25;
  const int m = act ? tid : 0;
  const long long i = sm_ut[m];
  const double pc0 = P->u_pcs[0LL * 650LL + i];
  const double pc1 = P->u_pcs[1LL * 650LL + i];
  const int f1 = sm_ut[50 + m];
  const double pc2 = P->u_pcs[2LL * 650LL + i];
  const int f2 = sm_ut[75 + m];
  const double pc4 = P->u_pcs[3LL * 650LL + i];
  const double pc7 = P->u_pcs[4LL * 650LL + i];
  const double pc10 = P->u_pcs[5LL * 650LL + i];
  const int f10 = sm_ut[150 + m];
  const double pc13 = P->u_pcs[6LL * 650LL + i];
  const int f13 = sm_ut[175 + m];
  const double pc16 = P->u_pcs[7LL * 650LL + i];
  const double x6 = P->u_pay1[0LL * 650LL + i];
  const double x9 = P->u_pay1[1LL * 650LL + i];
  const double x12 = P->u_pay1[2LL * 650LL + i];
  const double x15 = P->u_pay1[3LL * 650LL + i];
  const double x18 = P->u_pay1[4LL * 650LL + i];
  const double x21 = P->u_pay1[5LL * 650LL + i];
  const double s3 = __ldcg(P->br_sig + 0LL * 47008LL + f1);
  const double s4 = __ldcg(P->br_sig + 1LL * 47008LL + f1);
  const double s5 = __ldcg(P->br_sig + 0LL * 47008LL + f2);
  const double s6 = __ldcg(P->br_sig + 1LL * 47008LL + f2);
  const double s7 = __ldcg(P->br_sig + 2LL * 47008LL + f2);
  const double s14 = __ldcg(P->br_sig + 0LL * 47008LL + f10);
  const double s15 = __ldcg(P->br_sig + 1LL * 47008LL + f10);
  const double s16 = __ldcg(P->br_sig + 2LL * 47008LL + f10);
  const double s17 = __ldcg(P->br_sig + 0LL * 47008LL + f13);
  const double s18 = __ldcg(P->br_sig + 1LL * 47008LL + f13);
  const double s19 = __ldcg(P->br_sig + 2LL * 47008LL + f13);
  const double qo_0 = 1.0;
  const double qo_1 = qo_0;
  const double qo_2 = qo_0;
  const double qo_4 = qo_1 * s4;
  const double qo_7 = qo_2 * s7;
  const double qo_10 = qo_4;
  const double qo_13 = qo_7;
  const double qo_16 = qo_10 * s16;
  const double v20 = sm_fan[m * 9 + 7];
  const double v21 = ((pc16 * qo_16)) * x21;
  const double v22 = sm_fan[m * 9 + 8];
  if (act) {
    sm_q[225 + m] = v20;
    sm_q[250 + m] = v21;
    sm_q[275 + m] = v22;
  }
  __syncthreads();
  double V16;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 225;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 250;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 275;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 2; } }
    V16 = v20;
    if (best == 1) V16 = v21;
    if (best == 2) V16 = v22;
    if (tid == 0) P->br_astar[sm_ut[233 + 3]] = best;
  }
  const double v17 = sm_fan[m * 9 + 5];
  const double v18 = ((pc13 * qo_13) * s18) * x18;
  const double v19 = sm_fan[m * 9 + 6];
  double V13 = 0.0;
  V13 = V13 + v17;
  V13 = V13 + v18;
  V13 = V13 + v19;
  const double v14 = sm_fan[m * 9 + 4];
  const double v15 = ((pc10 * qo_10) * s15) * x15;
  const double v16 = V16;
  double V10 = 0.0;
  V10 = V10 + v14;
  V10 = V10 + v15;
  V10 = V10 + v16;
  const double v11 = sm_fan[m * 9 + 3];
  const double v12 = ((pc7 * qo_7)) * x12;
  const double v13 = V13;
  if (act) {
    sm_q[150 + m] = v11;
    sm_q[175 + m] = v12;
    sm_q[200 + m] = v13;
  }
  __syncthreads();
  double V7;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 150;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 175;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 200;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 2; } }
    V7 = v11;
    if (best == 1) V7 = v12;
    if (best == 2) V7 = v13;
    if (tid == 0) P->br_astar[sm_ut[233 + 2]] = best;
  }
  const double v8 = sm_fan[m * 9 + 2];
  const double v9 = ((pc4 * qo_4)) * x9;
  const double v10 = V10;
  if (act) {
    sm_q[75 + m] = v8;
    sm_q[100 + m] = v9;
    sm_q[125 + m] = v10;
  }
  __syncthreads();
  double V4;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 75;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 100;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 125;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 2; } }
    V4 = v8;
    if (best == 1) V4 = v9;
    if (best == 2) V4 = v10;
    if (tid == 0) P->br_astar[sm_ut[233 + 1]] = best;
  }
  const double v5 = sm_fan[m * 9 + 1];
  const double v6 = ((pc2 * qo_2) * s6) * x6;
  const double v7 = V7;
  double V2 = 0.0;
  V2 = V2 + v5;
  V2 = V2 + v6;
  V2 = V2 + v7;
  const double v3 = sm_fan[m * 9 + 0];
  const double v4 = V4;
  double V1 = 0.0;
  V1 = V1 + v3;
  V1 = V1 + v4;
  const double v1 = V1;
  const double v2 = V2;
  if (act) {
    sm_q[0 + m] = v1;
    sm_q[25 + m] = v2;
  }
  __syncthreads();
  double V0;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 0;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 25;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    V0 = v1;
    if (best == 1) V0 = v2;
    if (tid == 0) P->br_astar[sm_ut[233 + 0]] = best;
  }
  if (act) P->br_top1[P->tile_root[i]] = V0;
  __syncthreads();
  (void)qo_0;
}
__device__ __forceinline__ void micro_br1(const FusedParams* __restrict__ P, long long unit, double* sm_q) {
  const int tix = threadIdx.x % TEAM;
  const int gl = tix / 24, k = tix - gl * 24;
  const long long g = unit * 10 + gl;
  const bool ok = gl < 10 && g < 5850LL;
  const int glr = gl < 10 ? gl : 0;          // column block read by the reductions (idle lanes stay in bounds)
  const long long slot = ok ? g * 24 + k : 0;
  const int pat = P->mem_pat1[slot], ri = P->mem_ridx1[slot], pi = P->mem_pidx1[slot];
  const double pcv = P->mem_pc1[slot];
  const int f1 = P->mem_opp1[0LL * 140400LL + slot];
  const int f2 = P->mem_opp1[1LL * 140400LL + slot];
  const int f5 = P->mem_opp1[2LL * 140400LL + slot];
  const int f6 = P->mem_opp1[3LL * 140400LL + slot];
  const int4* pp = reinterpret_cast<const int4*>(P->paytab8 + (long long)pat * 16);
  const int4 pw0 = __ldg(pp + 0);
  const double x3 = (double)(int)(signed char)(pw0.x >> 0);
  const double x5 = (double)(int)(signed char)(pw0.x >> 8);
  const double x6 = (double)(int)(signed char)(pw0.x >> 16);
  const double x8 = (double)(int)(signed char)(pw0.x >> 24);
  const double x9 = (double)(int)(signed char)(pw0.y >> 0);
  const double x11 = (double)(int)(signed char)(pw0.y >> 8);
  const double x12 = (double)(int)(signed char)(pw0.y >> 16);
  const double x14 = (double)(int)(signed char)(pw0.y >> 24);
  const double x15 = (double)(int)(signed char)(pw0.z >> 0);
  const double x17 = (double)(int)(signed char)(pw0.z >> 8);
  const double x18 = (double)(int)(signed char)(pw0.z >> 16);
  const double x19 = (double)(int)(signed char)(pw0.z >> 24);
  const double x20 = (double)(int)(signed char)(pw0.w >> 0);
  const double x21 = (double)(int)(signed char)(pw0.w >> 8);
  const double x22 = (double)(int)(signed char)(pw0.w >> 16);
  const double rho_0 = pcv * __ldcg(P->br_reach1 + ri);
  const double s3 = __ldcg(P->br_sig + 0LL * 47008LL + f1);
  const double s4 = __ldcg(P->br_sig + 1LL * 47008LL + f1);
  const double s5 = __ldcg(P->br_sig + 0LL * 47008LL + f2);
  const double s6 = __ldcg(P->br_sig + 1LL * 47008LL + f2);
  const double s7 = __ldcg(P->br_sig + 2LL * 47008LL + f2);
  const double s14 = __ldcg(P->br_sig + 0LL * 47008LL + f5);
  const double s15 = __ldcg(P->br_sig + 1LL * 47008LL + f5);
  const double s16 = __ldcg(P->br_sig + 2LL * 47008LL + f5);
  const double s17 = __ldcg(P->br_sig + 0LL * 47008LL + f6);
  const double s18 = __ldcg(P->br_sig + 1LL * 47008LL + f6);
  const double s19 = __ldcg(P->br_sig + 2LL * 47008LL + f6);
  const double rho_1 = rho_0;
  const double rho_2 = rho_0;
  const double rho_3 = rho_1 * s4;
  const double rho_4 = rho_2 * s7;
  const double rho_5 = rho_3;
  const double rho_6 = rho_4;
  const double rho_7 = rho_5 * s16;
  const double v20 = (rho_7) * -x20;
  const double v21 = (rho_7) * -x21;
  const double v22 = (rho_7) * -x22;
  if (ok) {
    sm_q[9 * SMS + tix] = v20;
    sm_q[10 * SMS + tix] = v21;
    sm_q[11 * SMS + tix] = v22;
  }
  TEAM_SYNC();
  double V7;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 9 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 10 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 11 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 2; } }
    V7 = v20;
    if (best == 1) V7 = v21;
    if (best == 2) V7 = v22;
    if (ok && k == 0) P->br_astar[P->grp_iset1[g * 4 + 3]] = best;
  }
  const double v17 = (rho_6 * s17) * -x17;
  const double v18 = (rho_6 * s18) * -x18;
  const double v19 = (rho_6 * s19) * -x19;
  double V6 = 0.0;
  V6 = V6 + v17;
  V6 = V6 + v18;
  V6 = V6 + v19;
  const double v14 = (rho_5 * s14) * -x14;
  const double v15 = (rho_5 * s15) * -x15;
  const double v16 = V7;
  double V5 = 0.0;
  V5 = V5 + v14;
  V5 = V5 + v15;
  V5 = V5 + v16;
  const double v11 = (rho_4) * -x11;
  const double v12 = (rho_4) * -x12;
  const double v13 = V6;
  if (ok) {
    sm_q[6 * SMS + tix] = v11;
    sm_q[7 * SMS + tix] = v12;
    sm_q[8 * SMS + tix] = v13;
  }
  TEAM_SYNC();
  double V4;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 6 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 7 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 8 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 2; } }
    V4 = v11;
    if (best == 1) V4 = v12;
    if (best == 2) V4 = v13;
    if (ok && k == 0) P->br_astar[P->grp_iset1[g * 4 + 2]] = best;
  }
  const double v8 = (rho_3) * -x8;
  const double v9 = (rho_3) * -x9;
  const double v10 = V5;
  if (ok) {
    sm_q[3 * SMS + tix] = v8;
    sm_q[4 * SMS + tix] = v9;
    sm_q[5 * SMS + tix] = v10;
  }
  TEAM_SYNC();
  double V3;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 3 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 4 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 5 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 2; } }
    V3 = v8;
    if (best == 1) V3 = v9;
    if (best == 2) V3 = v10;
    if (ok && k == 0) P->br_astar[P->grp_iset1[g * 4 + 1]] = best;
  }
  const double v5 = (rho_2 * s5) * -x5;
  const double v6 = (rho_2 * s6) * -x6;
  const double v7 = V4;
  double V2 = 0.0;
  V2 = V2 + v5;
  V2 = V2 + v6;
  V2 = V2 + v7;
  const double v3 = (rho_1 * s3) * -x3;
  const double v4 = V3;
  double V1 = 0.0;
  V1 = V1 + v3;
  V1 = V1 + v4;
  const double v1 = V1;
  const double v2 = V2;
  if (ok) {
    sm_q[0 * SMS + tix] = v1;
    sm_q[1 * SMS + tix] = v2;
  }
  TEAM_SYNC();
  double V0;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 0 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 1 * SMS + glr * 24;
      for (int m = 0; m < 24; ++m) q = q + col[m];
      if (q > bq) { bq = q; best = 1; } }
    V0 = v1;
    if (best == 1) V0 = v2;
    if (ok && k == 0) P->br_astar[P->grp_iset1[g * 4 + 0]] = best;
  }
  if (ok) P->br_pv1[pi] = V0;
  TEAM_SYNC();
}
__device__ __noinline__ void upper_br1(const FusedParams* __restrict__ P, int u, double* sm_fan, double* sm_q, const int* sm_ut) {
  const int tid = threadIdx.x;
  const double* pvl = P->br_pv1 + 140400LL + (long long)u * 5400LL;
  const int* fd = P->fan_desc;
  for (int x = tid; x < 225; x += BLOCK) {
    const int mm = x / 9, f = x - mm * 9;
    const int k0 = fd[4 * f], nch = fd[4 * f + 1];
    const double* pv = pvl + (long long)k0 * 25 + mm;
    double acc = 0.0;
    double vv[24];
#pragma unroll
    for (int ch = 0; ch < 24; ++ch) vv[ch] = __ldcg(pv + ch * 25);
#pragma unroll
    for (int ch = 0; ch < 24; ++ch) acc = acc + vv[ch];
    (void)nch;
    sm_fan[x] = acc;
  }
  __syncthreads();
  const bool act = tid < 25;
  const int m = act ? tid : 0;
  const long long i = sm_ut[m];
  const double pc0 = P->u_pcs[0LL * 650LL + i];
  const int f0 = sm_ut[25 + m];
  const double pc1 = P->u_pcs[1LL * 650LL + i];
  const double pc2 = P->u_pcs[2LL * 650LL + i];
  const double pc4 = P->u_pcs[3LL * 650LL + i];
  const int f4 = sm_ut[100 + m];
  const double pc7 = P->u_pcs[4LL * 650LL + i];
  const int f7 = sm_ut[125 + m];
  const double pc10 = P->u_pcs[5LL * 650LL + i];
  const double pc13 = P->u_pcs[6LL * 650LL + i];
  const double pc16 = P->u_pcs[7LL * 650LL + i];
  const int f16 = sm_ut[200 + m];
  const double x6 = P->u_pay1[0LL * 650LL + i];
  const double x9 = P->u_pay1[1LL * 650LL + i];
  const double x12 = P->u_pay1[2LL * 650LL + i];
  const double x15 = P->u_pay1[3LL * 650LL + i];
  const double x18 = P->u_pay1[4LL * 650LL + i];
  const double x21 = P->u_pay1[5LL * 650LL + i];
  const double s1 = __ldcg(P->br_sig + 0LL * 47008LL + f0);
  const double s2 = __ldcg(P->br_sig + 1LL * 47008LL + f0);
  const double s8 = __ldcg(P->br_sig + 0LL * 47008LL + f4);
  const double s9 = __ldcg(P->br_sig + 1LL * 47008LL + f4);
  const double s10 = __ldcg(P->br_sig + 2LL * 47008LL + f4);
  const double s11 = __ldcg(P->br_sig + 0LL * 47008LL + f7);
  const double s12 = __ldcg(P->br_sig + 1LL * 47008LL + f7);
  const double s13 = __ldcg(P->br_sig + 2LL * 47008LL + f7);
  const double s20 = __ldcg(P->br_sig + 0LL * 47008LL + f16);
  const double s21 = __ldcg(P->br_sig + 1LL * 47008LL + f16);
  const double s22 = __ldcg(P->br_sig + 2LL * 47008LL + f16);
  const double qo_0 = 1.0;
  const double qo_1 = qo_0 * s1;
  const double qo_2 = qo_0 * s2;
  const double qo_4 = qo_1;
  const double qo_7 = qo_2;
  const double qo_10 = qo_4 * s10;
  const double qo_13 = qo_7 * s13;
  const double qo_16 = qo_10;
  const double v20 = sm_fan[m * 9 + 7];
  const double v21 = ((pc16 * qo_16) * s21) * -x21;
  const double v22 = sm_fan[m * 9 + 8];
  double V16 = 0.0;
  V16 = V16 + v20;
  V16 = V16 + v21;
  V16 = V16 + v22;
  const double v17 = sm_fan[m * 9 + 5];
  const double v18 = ((pc13 * qo_13)) * -x18;
  const double v19 = sm_fan[m * 9 + 6];
  if (act) {
    sm_q[225 + m] = v17;
    sm_q[250 + m] = v18;
    sm_q[275 + m] = v19;
  }
  __syncthreads();
  double V13;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 225;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 250;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 275;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 2; } }
    V13 = v17;
    if (best == 1) V13 = v18;
    if (best == 2) V13 = v19;
    if (tid == 0) P->br_astar[sm_ut[233 + 3]] = best;
  }
  const double v14 = sm_fan[m * 9 + 4];
  const double v15 = ((pc10 * qo_10)) * -x15;
  const double v16 = V16;
  if (act) {
    sm_q[150 + m] = v14;
    sm_q[175 + m] = v15;
    sm_q[200 + m] = v16;
  }
  __syncthreads();
  double V10;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 150;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 175;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 200;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 2; } }
    V10 = v14;
    if (best == 1) V10 = v15;
    if (best == 2) V10 = v16;
    if (tid == 0) P->br_astar[sm_ut[233 + 2]] = best;
  }
  const double v11 = sm_fan[m * 9 + 3];
  const double v12 = ((pc7 * qo_7) * s12) * -x12;
  const double v13 = V13;
  double V7 = 0.0;
  V7 = V7 + v11;
  V7 = V7 + v12;
  V7 = V7 + v13;
  const double v8 = sm_fan[m * 9 + 2];
  const double v9 = ((pc4 * qo_4) * s9) * -x9;
  const double v10 = V10;
  double V4 = 0.0;
  V4 = V4 + v8;
  V4 = V4 + v9;
  V4 = V4 + v10;
  const double v5 = sm_fan[m * 9 + 1];
  const double v6 = ((pc2 * qo_2)) * -x6;
  const double v7 = V7;
  if (act) {
    sm_q[75 + m] = v5;
    sm_q[100 + m] = v6;
    sm_q[125 + m] = v7;
  }
  __syncthreads();
  double V2;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 75;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 100;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    { double q = 0.0; const double* col = sm_q + 125;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 2; } }
    V2 = v5;
    if (best == 1) V2 = v6;
    if (best == 2) V2 = v7;
    if (tid == 0) P->br_astar[sm_ut[233 + 1]] = best;
  }
  const double v3 = sm_fan[m * 9 + 0];
  const double v4 = V4;
  if (act) {
    sm_q[0 + m] = v3;
    sm_q[25 + m] = v4;
  }
  __syncthreads();
  double V1;
  {
    int best = 0; double bq = 0.0;
    { double q = 0.0; const double* col = sm_q + 0;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (true) { bq = q; best = 0; } }
    { double q = 0.0; const double* col = sm_q + 25;
      for (int mm = 0; mm < 25; ++mm) q = q + col[mm];
      if (q > bq) { bq = q; best = 1; } }
    V1 = v3;
    if (best == 1) V1 = v4;
    if (tid == 0) P->br_astar[sm_ut[233 + 0]] = best;
  }
  const double v1 = V1;
  const double v2 = V2;
  double V0 = 0.0;
  V0 = V0 + v1;
  V0 = V0 + v2;
  if (act) P->br_top2[P->tile_root[i]] = V0;
  __syncthreads();
  (void)qo_0;
}
extern "C" __global__ void __launch_bounds__(BLOCK, 3) br_fused_k(const __grid_constant__ FusedParams PV) {
  const FusedParams* __restrict__ P = &PV;
  __shared__ double sm_q[3084];
  __shared__ double sm_fan[225];
  __shared__ double sm_s[11];
  __shared__ int sm_ut[237];
  __shared__ int s_lead;
  const unsigned int nb = gridDim.x;
  unsigned int nbar = 0;
  const int tid = threadIdx.x;
  const int NU0 = 26, NU1 = 26;
  if (tid == 0) {
    unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned int rank = atomicAdd(P->bar + 32 + 32 * nb + (smid & 1023u), 1u);
    s_lead = rank == 0 ? (int)atomicAdd(P->bar + 1, 1u) : -1;
  }
  __syncthreads();
  const int lead = s_lead;
  // phase 0: both players' reach to the portals under the evaluated strategy
  for (int u = blockIdx.x; u < NU0 + NU1; u += nb) {
    if (u < NU0) {
      for (int x = tid; x < 11; x += BLOCK) sm_s[x] = __ldcg(P->br_sig + (long long)UCA0[x] * 47008LL + P->tg_fid0[u * 4 + UCJ0[x]]);
      __syncthreads();
      if (tid == 0) upper_reach0(P->br_reach1, u, sm_s);
    } else {
      const int uu = u - NU0;
      for (int x = tid; x < 11; x += BLOCK) sm_s[x] = __ldcg(P->br_sig + (long long)UCA1[x] * 47008LL + P->tg_fid1[uu * 4 + UCJ1[x]]);
      __syncthreads();
      if (tid == 0) upper_reach1(P->br_reach2, uu, sm_s);
    }
    __syncthreads();
  }
  grid_barrier(P->bar, ++nbar, nb);
  const int n_lead = (int)__ldcg(P->bar + 1);
  // phase 1: the micro groups of both responders
  {
    const int team = tid / TEAM;
    double* t_q = sm_q + team * 3084;
    const long long nu0 = 585LL, nu1 = 585LL;
    for (long long un = (long long)blockIdx.x * 1 + team; un < nu0 + nu1; un += (long long)nb * 1) {
      if (un < nu0) micro_br0(P, un, t_q); else micro_br1(P, un - nu0, t_q);
    }
  }
  grid_barrier(P->bar, ++nbar, nb);
  // phase 2: the upper groups of both responders, one leader CTA per SM
  if (lead >= 0) {
    for (int u = lead; u < NU0 + NU1; u += n_lead) {
      if (u < NU0) upper_cache0(P, u, sm_ut); else upper_cache1(P, u - NU0, sm_ut);
      __syncthreads();
      if (u < NU0) upper_br0(P, u, sm_fan, sm_q, sm_ut); else upper_br1(P, u - NU0, sm_fan, sm_q, sm_ut);
    }
  }
  grid_barrier(P->bar, ++nbar, nb);
  // phase 3: the all-chance trunk above the tiles, children in order (one thread per responder)
  if (blockIdx.x < 2) {
    double* val = blockIdx.x == 0 ? P->br_top1 : P->br_top2;
    const long long* lv = &P->tl0;
    for (int l = (int)P->n_tl - 1; l >= 0; --l) {          // one thread per trunk node of the level
      const long long hi = lv[l + 1] < P->n_top ? lv[l + 1] : P->n_top;
      for (long long node = lv[l] + tid; node < hi; node += BLOCK) {
        if (P->g_player[node] != 0) continue;
        double acc = 0.0;
        const int c0 = P->g_first_child[node], c1 = P->g_first_child[node + 1];
        int ch = c0;
        for (; ch + 8 <= c1; ch += 8) {
          double vv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) vv[q] = __ldcg(val + ch + q);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc = acc + vv[q];
        }
        for (; ch < c1; ++ch) acc = acc + __ldcg(val + ch);
        val[node] = acc;
      }
      __syncthreads();
    }
    if (tid == 0) P->br_out[blockIdx.x] = __ldcg(val);
  }
}
