// rlm_kernels.cu -- the fused tick + learner-step kernel (sm_100a) and its launch wrappers.
//
// Execution model: one warp owns one environment for the whole launch.  The env record is
// staged HBM -> shared memory once, `n_ticks` market ticks are processed, and the record is
// written back.  Per tick:
//   lane 0      produces / fetches the tick message and runs the scalar market logic
//               (Intraday::NextState, src/environment/intraday.cpp:224-272);
//   lanes 0..7  push the eight rolling windows in parallel (one window per lane);
//   whenever the midprice has moved (Base::performAction's do-while, base.cpp:285-305) the
//   whole warp runs the learner step (serial.cpp:64-65): lane j hashes tiling j
//   (N_TILINGS == 32 == warp width), gathers theta, and the exact-order Q sums, the fused
//   trace-decay/clear/set/theta-update pass and the action selection follow.
#include <cuda_runtime.h>
#include <stdint.h>
#define RLM_TABLE_QUAL static __device__ const
#include "rlm_flow_tables.h"
#include "rlm_rndseq.h"
#include "rlm_agent.cuh"
#include "rlm_kernels.h"

// ---- shared memory carve-up -----------------------------------------------------------------
// [rndseq 8192][skellam 4096][pois30 4096][pois1p5 256] then per warp:
// [EnvHdr+rings : env_stride][scratch : SCRATCH_BYTES]
#define TABLE_BYTES (8192 + 4096 + 4096 + 256)
// per-warp scratch: [vbuf: (1 or 2) * A_max * VROW doubles][msg 128][pushv 10 doubles][to_vars 16 floats]
//                   [q_pre_a, q_pre_b: 18 doubles][small set: 64 ints]
#define SCR_MSG 0
#define SCR_PUSH (SCR_MSG + 128)
#define SCR_VARS (SCR_PUSH + 8 * RLM_NWIN)
#define SCR_Q (SCR_VARS + 64)
#define SCR_SS (SCR_Q + 8 * 2 * RLM_MAX_ACTIONS)
#define SCR_VBUF (SCR_SS + 4 * SS_SLOTS)
size_t rlm_scratch_bytes(int is_double) {
  return ((size_t)SCR_VBUF + (size_t)(is_double ? 2 : 1) * RLM_MAX_ACTIONS * VROW * 8 + 15) & ~(size_t)15;
}
size_t rlm_smem_bytes(int warps_per_cta, int env_stride, int scratch_bytes) {
  return TABLE_BYTES + (size_t)warps_per_cta * ((size_t)env_stride + (size_t)scratch_bytes);
}

cudaError_t rlm_upload_params(const DevParams* p) { return cudaMemcpyToSymbol(P, p, sizeof(DevParams)); }

// ---------------------------------------------------------------------------------------------
// init: Intraday ctor/Initialise state + RNG seeding, one thread per env.
// mode 0: full create; mode 1: episode reset (Base::Initialise base.cpp:123-135 keeps window sums, A13)
__global__ void rlm_init_kernel(DevPtrs ptr, int mode) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.n_envs) return;
  EnvHdr* e = (EnvHdr*)(ptr.env + (size_t)b * P.env_stride);
  double* ring = (double*)((unsigned char*)e + sizeof(EnvHdr));
  if (mode == 0) {
    unsigned char* raw = (unsigned char*)e;
    for (int i = 0; i < P.env_stride; ++i) raw[i] = 0;
    for (int i = 0; i < P.ring_total; ++i) ring[i] = 0.0;
    e->tp_val = -1.0;  // TargetPrice::val_ (target_price.cpp:8-10)
  }
  side_reset(e->side[0]);
  side_reset(e->side[1]);
  e->ask_quote = 0.0; e->bid_quote = 0.0;
  e->ep_reward = 0.0; e->ep_pnl = 0.0; e->ep_bandh = 0.0;
  e->market_buys = 0; e->market_sells = 0;
  e->ts_total = e->ts_ask = e->ts_bid = e->ts_both = e->ts_pos = e->ts_long = e->ts_short = 0;
  for (int w = 0; w < RLM_NWIN; ++w) { e->w_head[w] = 0; e->w_count[w] = 0; }
  e->last_date = 0; e->date = 0; e->time_ms = 0;
  e->phase = PH_PREOPEN;
  e->ep_step = 0;
  e->null_from = 1;
  e->stream_pos = 0;
  rlm_flow_init(&e->flow, &P.flow, (uint64_t)(P.env_index0 + b));
}

__global__ void rlm_seed_kernel(DevPtrs ptr, unsigned random_seed) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.n_envs) return;
  EnvHdr* e = (EnvHdr*)(ptr.env + (size_t)b * P.env_stride);
  unsigned seed = random_seed + (unsigned)(P.env_index0 + b);
  // std::mt19937_64::seed(value)
  for (int g = 0; g < 2; ++g) {
    unsigned long long* x = g == 0 ? ptr.mt_pol : ptr.mt_agt;
    if (!x) continue;
    x += (size_t)b * 312;
    unsigned long long v = (unsigned long long)seed;
    x[0] = v;
    for (int i = 1; i < 312; ++i) { v = 6364136223846793005ull * (v ^ (v >> 62)) + (unsigned long long)i; x[i] = v; }
  }
  e->mt_pol_idx = 312;
  e->mt_agt_idx = 312;
  // glibc srandom_r, TYPE_3
  unsigned s = seed == 0 ? 1u : seed;
  e->crand_r[0] = (int)s;
  for (int i = 1; i < 31; ++i) {
    long long hi = e->crand_r[i - 1] / 127773, lo = e->crand_r[i - 1] % 127773;
    long long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    e->crand_r[i] = (int)word;
  }
  e->crand_f = 3; e->crand_b = 0;
  for (int i = 0; i < 310; ++i) crand_next(*e);
}

// theta[i] = 2*U(0,1)-1 from the agent generator (agent.cpp:37-39,190-192), one thread per policy
__global__ void rlm_random_init_kernel(DevPtrs ptr, int n_policies) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_policies) return;
  EnvHdr* e = (EnvHdr*)(ptr.env + (size_t)b * P.env_stride);
  unsigned long long* x = ptr.mt_agt + (size_t)b * 312;
  double* th = ptr.theta + (size_t)b * P.memory_size;
  for (long long i = 0; i < P.memory_size; ++i) th[i] = 2.0 * mt_uniform_real(x, e->mt_agt_idx) - 1.0;
  if (ptr.theta_b) {
    double* tb = ptr.theta_b + (size_t)b * P.memory_size;
    for (long long i = 0; i < P.memory_size; ++i) tb[i] = 2.0 * mt_uniform_real(x, e->mt_agt_idx) - 1.0;
  }
}

// Agent::HandleTerminal's traces.decay(0.0) (agent.cpp:105)
__global__ void rlm_clear_traces_kernel(DevPtrs ptr) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.n_envs) return;
  EnvHdr* e = (EnvHdr*)(ptr.env + (size_t)b * P.env_stride);
  e->n_traces = 0;
}

// ---------------------------------------------------------------------------------------------
// parity record (include/rlm_record.h); lane 0 fills everything but the trace hash
__device__ __noinline__ void fill_record(rlm_step_record* r, const EnvHdr& e, const float* to_vars, unsigned long long thash) {
  r->step = e.ep_step; r->action = e.cur_action; r->time_ms = e.time_ms; r->terminal = is_terminal(e) ? 1 : 0;
  r->position = e.position; r->ask_quote = e.ask_quote; r->bid_quote = e.bid_quote;
  r->ask_level = e.ask_level; r->bid_level = e.bid_level;
  r->reward = e.last_reward; r->pnl_step = e.pnl_step;
  r->ep_pnl = e.ep_pnl; r->ep_reward = e.ep_reward; r->ep_bandh = e.ep_bandh;
  for (int s = 0; s < 2; ++s) {
    rlm_order_rec& o = s == 0 ? r->ask : r->bid;
    const OrderD& d = e.side[s].ord;
    o.exists = d.live ? 1 : 0; o.pad = 0;
    o.price = d.live ? d.price : 0.0; o.q_head = d.live ? d.q_head : 0; o.q_tail = d.live ? d.q_tail : 0;
    o.executed = d.live ? (d.size - ord_remaining(d)) : 0;
  }
  r->ask_transactions = e.side[0].n_transacted; r->bid_transactions = e.side[1].n_transacted;
  r->market_buys = e.market_buys; r->market_sells = e.market_sells;
  r->lo_vol_step = e.lo_vol_step;
  r->n_state = P.n_state_vars;
  for (int i = 0; i < RLM_N_STATE_MAX + 1; ++i) r->state[i] = (i < P.n_state_vars) ? to_vars[i] : 0.0f;
  r->delta = e.last_delta;
  r->n_traces = e.n_traces; r->pad = 0;
  r->trace_hash = thash;
}

// Learner::_step up to the first NextState of performAction (serial.cpp:55-61, base.cpp:254-284);
// lane 0.  Needs q_from / qb_from.  Returns false when the episode is over.
__device__ __noinline__ bool begin_step(EnvHdr& e, unsigned long long* mt_pol, const DynParams& D) {
  if (is_terminal(e)) {
    clear_inventory(e);  // Runner::RunEpisode, serial.cpp:31
    e.phase = PH_DONE;
    return false;
  }
  int a = policy_action(e, e.q_from, e.qb_from, mt_pol, D);
  e.cur_action = a;
  e.last_action = a;
  e.lo_vol_step = 0;
  e.pnl_step = 0.0;
  e.momentum_pnl_step = 0.0;
  do_action(e, a);
  check_orders(e);
  update_stats(e);
  e.agg_r = get_reward(e);
  e.agg_pnl = e.pnl_step;
  e.agg_mpm = 0.0;
  return true;
}

__device__ __noinline__ void flow_next_dev(rlm_flow_state* s, const int8_t* sk, const uint8_t* p30, const uint8_t* p15, rlm_tick_msg* m) {
  rlm_flow_next(s, &P.flow, sk, p30, p15, m);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS <= 14 ? 2 : 1)) rlm_tick_kernel(DevPtrs ptr, DynParams D) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned* s_rnd = (unsigned*)smem;
  int8_t* s_skellam = (int8_t*)(smem + 8192);
  uint8_t* s_pois30 = (uint8_t*)(smem + 8192 + 4096);
  uint8_t* s_pois1p5 = (uint8_t*)(smem + 8192 + 8192);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 2048; i += WARPS * 32) s_rnd[i] = rlm_rndseq_table[i];
  for (int i = tid; i < 4096; i += WARPS * 32) { s_skellam[i] = rlm_flow_skellam20_lut[i]; s_pois30[i] = rlm_flow_pois30_lut[i]; }
  for (int i = tid; i < 256; i += WARPS * 32) s_pois1p5[i] = rlm_flow_pois1p5_lut[i];
  __syncthreads();

  const int env_raw = blockIdx.x * WARPS + warp;
  const bool active = env_raw < P.n_envs;
  const int env = active ? env_raw : 0;
  const int stride = P.env_stride;
  unsigned char* wbase = smem + TABLE_BYTES + (size_t)warp * (stride + P.scratch_bytes);
  EnvHdr& e = *(EnvHdr*)wbase;
  double* ring = (double*)(wbase + sizeof(EnvHdr));
  unsigned char* scratch = wbase + stride;
  rlm_tick_msg& msg = *(rlm_tick_msg*)(scratch + SCR_MSG);
  double* pushv = (double*)(scratch + SCR_PUSH);
  float* to_vars = (float*)(scratch + SCR_VARS);
  double* q_pre_a = (double*)(scratch + SCR_Q);
  double* q_pre_b = q_pre_a + RLM_MAX_ACTIONS;
  int* sset = (int*)(scratch + SCR_SS);
  double* vbuf = (double*)(scratch + SCR_VBUF);
  if (!active) {  // tail warps only keep the CTA barriers balanced
    if (D.tick_sync) for (int t = 0; t < D.n_ticks; ++t) __syncthreads();
    return;
  }

  // ---- stage the env record: coalesced 16-byte loads
  {
    const int4* src = (const int4*)(ptr.env + (size_t)env * stride);
    int4* dst = (int4*)wbase;
    for (int i = lane; i < stride / 16; i += 32) dst[i] = src[i];
  }
  __syncwarp();

  const size_t pol = P.shared_policy ? 0 : (size_t)env;
  double* theta_a = ptr.theta + pol * (size_t)P.memory_size;
  double* theta_b = ptr.theta_b ? ptr.theta_b + pol * (size_t)P.memory_size : nullptr;
  int* tf = ptr.trace_f + (size_t)env * P.trace_cap;
  float* te = ptr.trace_e + (size_t)env * P.trace_cap;
  unsigned long long* mt_pol = ptr.mt_pol + (size_t)env * 312;
  unsigned long long* mt_agt = ptr.mt_agt ? ptr.mt_agt + (size_t)env * 312 : nullptr;
  const int A = P.n_actions;
  unsigned long long n_ticks_done = 0, n_steps_done = 0, sum_z = 0;
  unsigned long long bases[3] = {0ull, 0ull, 0ull};
  bool stopped = false;

#pragma unroll 1
  for (int t = 0; t < D.n_ticks; ++t) {
    if (D.tick_sync) __syncthreads();
    const int phase = e.phase;
    if (phase == PH_DONE || stopped) continue;
    // ---- tick message
    if (P.source == RLM_SOURCE_GENERATOR) {
      if (lane == 0) flow_next_dev(&e.flow, s_skellam, s_pois30, s_pois1p5, &msg);
    } else {
      const int pos = D.stream_off + t;  // tick-synchronous: every env consumes the same tick index
      if (pos >= D.stream_ticks) {
        if (lane == 0) e.err |= ERR_STREAM_UNDERRUN;
        stopped = true;
        continue;
      }
      const unsigned* src = (const unsigned*)(ptr.stream + ((size_t)pos * P.n_envs + env));
      ((unsigned*)&msg)[lane] = __ldg(src + lane);  // one 128-byte line per tick
    }
    __syncwarp();

    if (phase == PH_PREOPEN) {  // intraday.cpp:111-116
      if (lane == 0) {
        rlm_tick_msg none = msg;
        none.n_tx = 0;
        update_book_profiles(e, none);
        if (market_is_open(e)) e.phase = PH_WARMUP;
      }
      __syncwarp();
      continue;
    }

    // ---- Intraday::NextState
    if (lane == 0) {
      if (phase == PH_RUN) e.pnl_step = 0.0;  // base.cpp:286
      next_state_scalar(e, msg, pushv);
    }
    __syncwarp();
    if (lane < 8) window_push(e, ring, lane, pushv[lane]);
    __syncwarp();
    n_ticks_done++;

    bool step_end = false;
    if (phase == PH_WARMUP) {  // intraday.cpp:118-135
      bool full = (lane < 8) ? (e.w_count[lane] == P.win_size[lane]) : true;
      full = __all_sync(FULL, full);
      if (lane == 0) {
        e.tp_val = e.w_mean[W_TP];
        if (full) {
          place_orders(e, 1, 1);
          e.null_from = 1;
          e.phase = PH_RUN;
        }
      }
      __syncwarp();
      if (!full) continue;
      // serial.cpp:24-25,55-60: the first from-state is the never-populated State (all features 0)
      double qa, qb;
      eval_q(s_rnd, theta_a, theta_b, e.from_vars, P.n_state_vars, true, vbuf, lane, qa, qb, bases, false);
      if (lane < A) { e.q_from[lane] = qa; e.qb_from[lane] = qb; }
      __syncwarp();
      if (lane == 0) begin_step(e, mt_pol, D);
      __syncwarp();
      continue;
    }

    // ---- PH_RUN: tail of one iteration of performAction's do-while (base.cpp:292-305)
    if (lane == 0) {
      e.tp_val = e.w_mean[W_TP];
      double mpm = m_midprice(e) - m_last_midprice(e);
      e.pnl_step += (double)e.position * mpm;
      e.momentum_pnl_step += (double)e.position * mpm;
      e.agg_r += get_reward(e);
      e.agg_pnl += e.pnl_step;
      e.agg_mpm += mpm;
      bool ex = !(!is_terminal(e) && fabs(e.agg_mpm) < 1e-5);
      if (ex) {  // base.cpp:317-331
        e.pnl_step = e.agg_pnl;
        pushv[W_PNLUP] = fmax(0.0, e.pnl_step);
        pushv[W_PNLDN] = fabs(fmin(0.0, e.pnl_step));
        e.ep_reward += e.agg_r;
        e.ep_bandh += e.agg_mpm;
      }
      pushv[0] = ex ? 1.0 : 0.0;
    }
    __syncwarp();
    step_end = pushv[0] != 0.0;
    if (!step_end) continue;

    // ================= learner step: serial.cpp:64-65 =================
    if (lane == W_PNLUP || lane == W_PNLDN) window_push(e, ring, lane, pushv[lane]);
    __syncwarp();
    if (lane == 0) {
      // State::newState -> Intraday::getState (state.cpp:35-43, intraday.cpp:411-416)
      for (int i = 0; i < P.n_state_vars; ++i) to_vars[i] = (float)get_variable(e, ring, P.state_vars[i]);
      e.last_reward = get_reward(e);
    }
    __syncwarp();
    // Q(to, .) under the current theta
    {
      double qa, qb;
      eval_q(s_rnd, theta_a, theta_b, to_vars, P.n_state_vars, false, vbuf, lane, qa, qb, bases, false);
      if (lane < A) { q_pre_a[lane] = qa; q_pre_b[lane] = qb; }
    }
    __syncwarp();
    // Agent::HandleTransition (agent.cpp:86-101): UpdateTraces decision + TD error; lane 0
    if (lane == 0) {
      const int action = e.cur_action;
      const double reward = e.last_reward;
      const double F_term = P.gamma * 0.0 - 0.0;  // potentials are 0 (base.cpp:239-242)
      float rate = P.gl;
      double delta;
      int table = 0;
      if (P.algorithm == RLM_ALGO_SARSA) {  // Agent::UpdateTraces :111-115, SARSA::UpdateWeights :300-311
        double Q1 = e.q_from[action];
        int a2 = policy_action(e, q_pre_a, q_pre_b, mt_pol, D);
        double Q2 = q_pre_a[a2];
        delta = reward + F_term + P.gamma * Q2 - Q1;
      } else if (P.algorithm == RLM_ALGO_Q_LEARN) {  // QLearn :272-292
        int amax = argmax_ties(e, e.q_from);
        if (action != amax) rate = 0.0f;
        double Q = e.q_from[action];
        int am2 = argmax_ties(e, q_pre_a);
        delta = reward + F_term + P.gamma * q_pre_a[am2] - Q;
      } else {  // DoubleQLearn :319-353
        int amax = argmax_ties(e, e.q_from);
        if (action != amax) rate = 0.0f;
        if (mt_uniform_real(mt_agt, e.mt_agt_idx) > 0.5) {
          double Qa = e.q_from[action];
          int am2 = argmax_ties(e, q_pre_a);
          delta = reward + F_term + P.gamma * q_pre_b[am2] - Qa;
          table = 0;
        } else {
          double Qb = e.qb_from[action];
          int am2 = argmax_ties(e, q_pre_b);
          delta = reward + F_term + P.gamma * q_pre_a[am2] - Qb;
          table = 1;
        }
      }
      e.last_delta = delta;
      pushv[0] = (double)rate;
      pushv[1] = (D.alpha * delta) / (double)RLM_N_TILINGS;  // Agent::updateQ: update / N_TILINGS
      pushv[2] = (double)table;
    }
    __syncwarp();
    {
      const float rate = (float)pushv[0];
      const double scaled = pushv[1];
      double* th = (pushv[2] != 0.0) ? theta_b : theta_a;
      int nz = trace_pass(e, sset, tf, te, th, e.cur_action, rate, scaled, lane);
      if (lane == 0) { e.n_traces = nz; e.sum_traces += nz; }
      sum_z += (lane == 0) ? (unsigned long long)nz : 0ull;
    }
    __syncwarp();
    __threadfence_block();
    // parity record
    if (env < P.record_envs) {
      unsigned long long h = trace_hash(tf, te, theta_a, e.n_traces, lane);
      if (lane == 0) {
        int c = ptr.record_count[env];
        if (c < P.record_cap) fill_record(&ptr.records[(size_t)env * P.record_cap + c], e, to_vars, h);
        ptr.record_count[env] = c + 1;
      }
    }
    // the to-state becomes the from-state; Q(from, .) under the UPDATED theta (serial.cpp:55,60)
    if (lane < RLM_N_STATE_MAX) e.from_vars[lane] = to_vars[lane];
    e.from_base0[lane] = mod_m(bases[0]);  // group-0 partial hash of the new from-state (trace_pass)
    if (lane == 0) { e.null_from = 0; e.n_steps++; e.ep_step++; }
    __syncwarp();
    {
      double qa, qb;
      eval_q(s_rnd, theta_a, theta_b, e.from_vars, P.n_state_vars, false, vbuf, lane, qa, qb, bases, true);
      if (lane < A) { e.q_from[lane] = qa; e.qb_from[lane] = qb; }
    }
    __syncwarp();
    n_steps_done++;
    if (lane == 0) begin_step(e, mt_pol, D);
    __syncwarp();
  }

  // ---- write the env record back and publish counters
  __syncwarp();
  {
    int4* dst = (int4*)(ptr.env + (size_t)env * stride);
    const int4* src = (const int4*)wbase;
    for (int i = lane; i < stride / 16; i += 32) dst[i] = src[i];
  }
  if (lane == 0) {
    atomicAdd(&ptr.counters[0], n_ticks_done);
    atomicAdd(&ptr.counters[1], n_steps_done);
    atomicAdd(&ptr.counters[2], sum_z);
    if (e.err) atomicOr(&ptr.counters[4], (unsigned long long)e.err);
  }
}

// ---------------------------------------------------------------------------------------------
template <int WARPS>
static cudaError_t launch_tick(const DevPtrs& ptr, const DynParams& D, int n_envs, int env_stride, int scratch_bytes, cudaStream_t st) {
  size_t smem = rlm_smem_bytes(WARPS, env_stride, scratch_bytes);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(rlm_tick_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  int grid = (n_envs + WARPS - 1) / WARPS;
  rlm_tick_kernel<WARPS><<<grid, WARPS * 32, smem, st>>>(ptr, D);
  return cudaGetLastError();
}

cudaError_t rlm_launch_tick(const DevPtrs& ptr, const DynParams& D, int n_envs, int env_stride, int scratch_bytes, int warps, cudaStream_t st) {
  switch (warps) {
    case 4: return launch_tick<4>(ptr, D, n_envs, env_stride, scratch_bytes, st);
    case 8: return launch_tick<8>(ptr, D, n_envs, env_stride, scratch_bytes, st);
    case 14: return launch_tick<14>(ptr, D, n_envs, env_stride, scratch_bytes, st);
    case 16: return launch_tick<16>(ptr, D, n_envs, env_stride, scratch_bytes, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t rlm_launch_init(const DevPtrs& ptr, int n_envs, int mode, cudaStream_t st) {
  rlm_init_kernel<<<(n_envs + 127) / 128, 128, 0, st>>>(ptr, mode);
  return cudaGetLastError();
}
cudaError_t rlm_launch_seed(const DevPtrs& ptr, int n_envs, unsigned seed, cudaStream_t st) {
  rlm_seed_kernel<<<(n_envs + 127) / 128, 128, 0, st>>>(ptr, seed);
  return cudaGetLastError();
}
cudaError_t rlm_launch_random_init(const DevPtrs& ptr, int n_policies, cudaStream_t st) {
  rlm_random_init_kernel<<<(n_policies + 63) / 64, 64, 0, st>>>(ptr, n_policies);
  return cudaGetLastError();
}
cudaError_t rlm_launch_clear_traces(const DevPtrs& ptr, int n_envs, cudaStream_t st) {
  rlm_clear_traces_kernel<<<(n_envs + 127) / 128, 128, 0, st>>>(ptr);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// unit-level device entry points (golden vectors of the reference's tests)
__global__ void k_test_to_ticks(const double* px, int n, int* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int err = 0;
  if (i < n) { int t = to_ticks(px[i], &err); out[i] = err ? -1 : t; }
}
__global__ void k_test_to_price(const int* ticks, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int err = 0;
  if (i < n) { double p = to_price(ticks[i], &err); out[i] = err ? -1.0 : p; }
}
// one warp per state: all tile indices, out[s][a][96]
__global__ void k_test_tiles(const float* vars, int n, int* out) {
  __shared__ unsigned s_rnd[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_rnd[i] = rlm_rndseq_table[i];
  __syncthreads();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* v = vars + (size_t)warp * P.n_state_vars;
  const int A = P.n_actions, nv = P.n_state_vars;
  for (int g = 0; g < 3; ++g) {
    const float* gv = (g == 1) ? v + 3 : v;
    const int nf = (g == 0) ? 3 : ((g == 1) ? nv - 3 : nv);
    unsigned long long base = tile_base_sum(s_rnd, gv, nf, lane);
    for (int a = 0; a < A; ++a) out[((size_t)warp * A + a) * 96 + g * 32 + lane] = tile_index(s_rnd, base, nf, g * A + a);
  }
}
__global__ void k_test_order(long long size, long long q_head, const rlm_order_op* ops, int n_ops, rlm_order_state* out) {
  if (threadIdx.x || blockIdx.x) return;
  OrderD o; o.live = 1; o.price = 1.0; o.size = size; o.q_head = q_head; o.q_tail = 0; o.executed = 0; o.initial_queue = q_head; o.transactions = 0;
  for (int i = 0; i < n_ops; ++i) {
    long long ret = 0;
    switch (ops[i].op) {
      case 0: ret = ord_do_transaction(o, ops[i].arg); break;
      case 1: ord_do_cancellation(o, ops[i].arg); break;
      case 2: o.q_tail += ops[i].arg; break;
      case 3: o.q_head = 0; o.q_tail = 0; break;
    }
    out[i].size = o.size; out[i].q_head = o.q_head; out[i].q_tail = o.q_tail; out[i].executed = o.executed; out[i].ret = ret;
  }
}
// RollingMean<double> through the production window_push, window slot W_MID
__global__ void k_test_rolling_mean(const double* vals, int n, double* out, double* ring_mem, EnvHdr* e) {
  if (threadIdx.x || blockIdx.x) return;
  for (int i = 0; i < n; ++i) {
    window_push(*e, ring_mem, W_MID, vals[i]);
    out[2 * i] = e->w_mean[W_MID];
    out[2 * i + 1] = e->w_s[W_MID] / (double)((unsigned long long)((long long)e->w_count[W_MID] - 1));
  }
}

cudaError_t rlm_launch_test_to_ticks(const double* px, int n, int* out) { k_test_to_ticks<<<(n + 127) / 128, 128>>>(px, n, out); return cudaGetLastError(); }
cudaError_t rlm_launch_test_to_price(const int* t, int n, double* out) { k_test_to_price<<<(n + 127) / 128, 128>>>(t, n, out); return cudaGetLastError(); }
cudaError_t rlm_launch_test_tiles(const float* vars, int n, int* out) { k_test_tiles<<<(n + 3) / 4, 128>>>(vars, n, out); return cudaGetLastError(); }
cudaError_t rlm_launch_test_order(long long size, long long q_head, const rlm_order_op* ops, int n_ops, rlm_order_state* out) { k_test_order<<<1, 32>>>(size, q_head, ops, n_ops, out); return cudaGetLastError(); }
cudaError_t rlm_launch_test_rolling_mean(const double* vals, int n, double* out, double* ring_mem, EnvHdr* e) { k_test_rolling_mean<<<1, 32>>>(vals, n, out, ring_mem, e); return cudaGetLastError(); }
