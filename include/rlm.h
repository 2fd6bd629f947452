/* rlm.h -- C ABI of the B200-native batched limit-order-book RL environment.
 *
 * Drop-in boundary for the hot path of tspooner/rl_markets.
 * The reference has no FFI; its seam is the shared library `rl_engine`
 * (src/CMakeLists.txt:5) and the C++ classes the driver uses by name.  Each
 * entry point below names the reference interface it replaces, batched over
 * `n_envs` independent environments:
 *
 *   rlm_create            environment::Intraday<>::Intraday(Config&)      include/environment/intraday.h:58
 *                         + rl::QLearn/SARSA/DoubleQLearn(policy, Config&) include/rl/agent.h:106-131
 *                         + rl::EpsilonGreedy/Greedy/Random(...)           include/rl/policy.h:31-66
 *   rlm_load_ticks        Intraday::LoadData / data::Streamer<R>           include/data/streamer.h:16-56
 *   rlm_reset             Intraday::Initialise                             src/environment/intraday.cpp:103-138
 *   rlm_run_ticks         experiment::serial::Learner::_step               src/experiment/serial.cpp:53-70
 *                         = Agent::action + Base::performAction + State::newState
 *                           + Agent::HandleTransition, repeated while ticks remain
 *   rlm_get_state         Intraday::getState                               src/environment/intraday.cpp:411-416
 *   rlm_get_reward        Base::getReward                                  src/environment/base.cpp:166-237
 *   rlm_handle_terminal   Agent::HandleTerminal + Policy::HandleTerminal   src/rl/agent.cpp:103-109, policy.cpp:79-82
 *   rlm_go_greedy         Agent::GoGreedy                                  src/rl/agent.cpp:76-79
 *   rlm_read_theta        Agent::write_theta (raw double[MEMORY_SIZE])     src/rl/agent.cpp:176-181
 *   rlm_get_stats         Base::getEpisodeReward/getEpisodePnL/...         src/environment/base.cpp:244-252,458-473
 *
 * Conventions: plain C types only; every call returns 0 on success or a
 * negative rlm_status; rlm_last_error() holds the message (the reference's
 * C++ exceptions never cross this boundary); handles are opaque; the library
 * owns all device memory; one host thread per handle.  There is NO CPU
 * fallback: rlm_create fails with RLM_ERR_NO_DEVICE when no CUDA device is
 * usable.
 */
#ifndef RLM_H
#define RLM_H

#include <stdint.h>
#include "rlm_flow.h"
#include "rlm_record.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RLM_ABI_VERSION 3

typedef enum rlm_status {
  RLM_OK = 0,
  RLM_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument in the reference (market.cpp:58,86,112,135) */
  RLM_ERR_RUNTIME = -2,          /* std::runtime_error   (book.cpp:74-77, order.cpp:22-27, ...)       */
  RLM_ERR_NO_DEVICE = -3,        /* CUDA device / extension unavailable: there is no CPU fallback     */
  RLM_ERR_CUDA = -4,
  RLM_ERR_UNSUPPORTED = -5,      /* configuration outside the B200 path (e.g. n_tilings != 32)        */
  RLM_ERR_END_OF_DATA = -6       /* stream exhausted (performAction returning false, base.cpp:289)    */
} rlm_status;

/* learning.algorithm (src/main.cpp:169-189) */
enum { RLM_ALGO_Q_LEARN = 0, RLM_ALGO_SARSA = 1, RLM_ALGO_DOUBLE_Q_LEARN = 2,
       RLM_ALGO_R_LEARN = 3, RLM_ALGO_ONLINE_R_LEARN = 4, RLM_ALGO_DOUBLE_R_LEARN = 5 };
/* policy.type (src/main.cpp:141-165) */
enum { RLM_POLICY_GREEDY = 0, RLM_POLICY_RANDOM = 1, RLM_POLICY_EPSILON_GREEDY = 2, RLM_POLICY_BOLTZMANN = 3 };
/* reward.measure (src/environment/base.cpp:55-75) */
enum { RLM_REWARD_NONE = 0, RLM_REWARD_PNL, RLM_REWARD_PNL_DAMPED, RLM_REWARD_SPREAD, RLM_REWARD_NORMED,
       RLM_REWARD_LOVOL, RLM_REWARD_MM_LINEAR, RLM_REWARD_MM_EXP, RLM_REWARD_MM_DIV };
/* state.variables (include/environment/intraday.h:17-23) */
enum { RLM_VAR_POS = 0, RLM_VAR_SPD, RLM_VAR_MPM, RLM_VAR_IMB, RLM_VAR_SVL, RLM_VAR_VOL, RLM_VAR_RSI,
       RLM_VAR_VWAP, RLM_VAR_A_DIST, RLM_VAR_A_QUEUE, RLM_VAR_B_DIST, RLM_VAR_B_QUEUE, RLM_VAR_LAST_ACTION };
/* market.target_price.type AS WRITTEN IN THE YAML.  The reference's selector is
 * inverted (src/environment/base.cpp:101-112): "midprice" instantiates
 * tp::MicroPrice, every other string instantiates tp::MidPrice; the string
 * "book" additionally switches the quote rule (intraday.cpp:64-71). */
enum { RLM_TP_YAML_MIDPRICE = 0, RLM_TP_YAML_MICROPRICE = 1, RLM_TP_YAML_VWAP = 2, RLM_TP_YAML_BOOK = 3 };
/* where ticks come from */
enum { RLM_SOURCE_GENERATOR = 0, /* rlm_flow.h generator evaluated inside the tick kernel */
       RLM_SOURCE_STREAM = 1 };  /* rlm_tick_msg chunks uploaded with rlm_load_ticks      */

#define RLM_MAX_BANDS 32 /* power of two (band search); the longest upstream table, NasdaqNordic / Oslo, has 17 bands */
#define RLM_MAX_ACTIONS 9
#define RLM_N_TILINGS 32

typedef struct rlm_config {
  /* ---- batch / placement (new; the reference is one env per thread) ---- */
  int32_t n_envs;        /* environments owned by this handle (this GPU)                           */
  int32_t device;        /* CUDA device ordinal                                                    */
  int64_t env_index0;    /* global index of local env 0; seeds and flow streams use env_index0 + b */
  int32_t shared_policy; /* 0: one theta per env (B reference processes); 1: one theta per handle  */
  int32_t source;        /* RLM_SOURCE_*                                                           */
  /* ---- learning.* (src/rl/agent.cpp:14-50, src/rl/state.cpp:22-24) ---- */
  int64_t memory_size;
  int32_t n_tilings;
  int32_t n_actions;
  int32_t algorithm;
  int32_t random_init;
  double group_weights[3];
  double gamma, lambda, omega, alpha_start, alpha_floor, beta;
  /* ---- policy.* (src/main.cpp:137-165; eps/tau are read as float there) ---- */
  int32_t policy_type;
  float eps_init, eps_floor;
  uint32_t eps_T;
  float tau_init, tau_floor;
  uint32_t tau_T;
  int32_t spread_lookback;
  /* ---- reward.* (src/environment/base.cpp:28-32,46-47,55-75) ---- */
  int32_t reward_measure;
  float damping_factor, pos_weight, trd_weight, pnl_weight;
  int32_t pnl_lookback;
  /* ---- state.* (src/environment/intraday.cpp:52-62, base.cpp:35-50) ---- */
  int32_t n_state_vars;
  int32_t state_vars[RLM_N_STATE_MAX];
  int32_t lb_mpm, lb_vlt, lb_svl, lb_rsi, lb_vwap;
  /* ---- market.* (base.cpp:18-25,101-112) ---- */
  int64_t pos_lb, pos_ub;
  int32_t order_size;
  int32_t target_price_type; /* RLM_TP_YAML_* */
  int32_t tp_lookback;
  /* ---- venue: Market::pts_ (price -> tick size), ascending (src/market/market.cpp:11-38,206-245) ---- */
  int32_t n_bands;
  double band_px[RLM_MAX_BANDS];
  double band_ts[RLM_MAX_BANDS];
  int64_t open_ms, close_ms; /* Market::mo_, mc_ */
  /* ---- debug.random_seed (main.cpp:84-88, agent.cpp:30): env b uses random_seed + env_index0 + b ---- */
  uint32_t random_seed;
  /* ---- synthetic flow (RLM_SOURCE_GENERATOR) ---- */
  rlm_flow_params flow;
  /* ---- capacity / debugging ---- */
  int32_t trace_cap;    /* max nonzero traces kept per env; 0 = derive from gamma*lambda (traces.h:13) */
  int32_t record_envs;  /* first N envs write one rlm_step_record per learner step (parity tests)    */
  int32_t record_cap;   /* records kept per recorded env                                              */
  int32_t reserved[5];
} rlm_config;

typedef struct rlm_handle_s* rlm_handle;

/* Aggregate counters since rlm_create (all envs of the handle). */
typedef struct rlm_counters {
  int64_t ticks;        /* NextState() calls                                   */
  int64_t steps;        /* completed learner steps (Learner::_step)             */
  int64_t sum_traces;   /* sum over steps of n_nonzero_traces after the update  */
  int64_t terminal_envs;/* envs currently terminal                              */
  int64_t kernel_launches;
} rlm_counters;

/* Per-env statistics (Base getters, src/environment/base.cpp:244-252,458-473). */
typedef struct rlm_env_stats {
  double episode_reward, episode_pnl, episode_bandh;
  int64_t position;
  int32_t ask_transactions, bid_transactions, market_buys, market_sells;
  int32_t total_ticks, steps;
  int32_t terminal, phase;
} rlm_env_stats;

const char* rlm_last_error(void);
int rlm_abi_version(void);

/* config/example.yaml defaults + LSE tick table for AAL (market.cpp:216-227). */
int rlm_config_default(rlm_config* cfg);

int rlm_create(const rlm_config* cfg, rlm_handle* out);
int rlm_destroy(rlm_handle h);

/* Intraday::Initialise for every env: books/windows cleared, stream rewound. */
int rlm_reset(rlm_handle h);

/* Which experiment::serial step the handle runs (src/experiment/serial.cpp):
 *   RLM_MODE_TRAIN    Learner::_step    :53-70   act on the previous state, performAction, newState, HandleTransition
 *   RLM_MODE_BACKTEST Backtester::_step :121-137 newState, act, performAction; theta and traces are never touched.
 * In backtest mode rlm_read_records yields one row per step with every column of Intraday::LogProfit's profit_log
 * (intraday.cpp:437-451) and rlm_get_stats the counters Base::writeStats dumps (base.cpp:451-456). */
enum { RLM_MODE_TRAIN = 0, RLM_MODE_BACKTEST = 1 };
int rlm_set_mode(rlm_handle h, int32_t mode);

/* `environment::Intraday<> env(c)` of src/main.cpp:219: every env object is rebuilt from scratch (window sums,
 * statistics, position, book, records) while the agents keep theta, traces, generator positions and schedules.
 * `flow` (may be NULL = keep) replaces the synthetic-flow parameters, i.e. "LoadData of another day". */
int rlm_new_env(rlm_handle h, const rlm_flow_params* flow);

/* RLM_SOURCE_STREAM: append `n_ticks` messages per env, host layout msgs[t][env] (tick-major).
 * The copy is issued on the handle's copy stream; the buffer must stay valid until rlm_sync. */
int rlm_load_ticks(rlm_handle h, const rlm_tick_msg* msgs, int32_t n_ticks);

/* Advance every env by n_ticks market ticks.  Each env runs warm-up, performAction's inner NextState loop, and --
 * whenever its midprice has moved -- the complete learner step.  On return every env has consumed exactly n_ticks
 * messages and sits inside performAction's loop, whatever order the engine ran the envs' ticks in (they never interact).
 * Calls shorter than 128 ticks only enqueue work (asynchronous; see rlm_sync); longer calls of independent policies on
 * up to 16 384 envs run round by round and return when the device is nearly done with them (the host follows the
 * device to learn when the last env has finished; RLM_ROUNDS=0 keeps every call asynchronous). */
int rlm_run_ticks(rlm_handle h, int32_t n_ticks);

int rlm_sync(rlm_handle h);

int rlm_get_counters(rlm_handle h, rlm_counters* out);
int rlm_get_stats(rlm_handle h, int32_t env0, int32_t n, rlm_env_stats* out);
int rlm_get_state(rlm_handle h, float* out /* [n_envs][n_state_vars] */);
int rlm_get_reward(rlm_handle h, double* out /* [n_envs] last reward handed to the agent */);
int rlm_get_actions(rlm_handle h, int32_t* out /* [n_envs] last action */);
/* rho of the R-learning agents (RLearn / OnlineRLearn / DoubleRLearn private member, include/rl/agent.h:131,145,157;
   the reference never prints it -- exposed here so that parity of the average-reward estimate can be checked) */
int rlm_get_rho(rlm_handle h, double* out /* [n_envs] */);
/* theta (both tables of a double agent) of every policy of `src` -> `dst`, device to device: the B-env form of handing one
   trained `Agent*` to the next phase (src/main.cpp:196-222 keeps `m` across training episodes and into evaluation).  The
   handles must agree in device, n_envs / shared_policy, memory_size and algorithm family; traces and env state of `dst`
   are untouched. */
int rlm_copy_theta(rlm_handle dst, rlm_handle src);
/* Diagnostic (no reference counterpart): number of weights of env b's table A that are not +0.0, i.e. how far the
   table has filled up (independent policies). */
int rlm_get_occupancy(rlm_handle h, int32_t* out /* [n_envs] */);

int rlm_handle_terminal(rlm_handle h, int32_t episode);
int rlm_go_greedy(rlm_handle h);

/* theta access: policy = env index (independent) or 0 (shared); table 0 = A, 1 = B (double agents). */
int rlm_read_theta(rlm_handle h, int32_t policy, int32_t table, double* out, int64_t n);
int rlm_write_theta(rlm_handle h, int32_t policy, int32_t table, const double* in, int64_t n);

/* parity dump of recorded envs (cfg.record_envs / record_cap) */
int rlm_read_records(rlm_handle h, int32_t env, rlm_step_record* out, int32_t cap, int32_t* n_out);

/* raw device pointers, for torch.distributed / NCCL plumbing on the shared-policy path */
int rlm_device_ptrs(rlm_handle h, void** theta, void** dtheta, int64_t* n_doubles);
/* Shared policy (cfg.shared_policy = 1; reference analogue: threads sharing one rl::Agent*,
 * src/main.cpp:196-206).  One training tick across GPUs is
 *   rlm_shared_tick_accumulate(h)   env tick + learner steps under theta_t, updates summed into dtheta
 *   all-reduce(dtheta, SUM)         caller's collective (torch.distributed / NCCL) on rlm_device_ptrs()
 *   rlm_apply_dtheta(h)             theta += dtheta; dtheta = 0; Q(from,.) under theta_{t+1}; next actions
 * On one GPU rlm_run_ticks does the same without the collective. */
int rlm_shared_tick_accumulate(rlm_handle h);
int rlm_apply_dtheta(rlm_handle h);
/* ---- split surface: the reference's Environment::step / Agent::update seam, batched ------------------------------
 * rlm_run_ticks fuses the whole of experiment::serial::Learner::_step (src/experiment/serial.cpp:53-70).  These three
 * calls expose its parts, so that an external policy can supply the actions and an external consumer can read every
 * transition; driven in the order below they reproduce rlm_run_ticks bit for bit (tests/test_gpu_split.py):
 *
 *   rlm_env_step(h, NULL, ..)        Runner::RunEpisode: environment.Initialise()              serial.cpp:18-25
 *   rlm_agent_update(h, NULL)        ... Q(first from-state, .) for the first action
 *   repeat:
 *     rlm_act(h, actions)            int action = m->action(*last_state)                       serial.cpp:60,  include/rl/agent.h:60
 *     rlm_env_step(h, actions, r, t) environment.performAction(action) + getReward()           serial.cpp:61,66, include/environment/base.h:132
 *     rlm_agent_update(h, delta)     state->newState(env); m->HandleTransition(...)            serial.cpp:64-67, include/rl/agent.h:62-67
 *
 * Every env advances by ONE learner step per rlm_env_step (its own K >= 1 market ticks, base.cpp:285-305); envs are
 * therefore not tick-aligned afterwards, which is why this surface needs source = generator.  actions_out[b] / the
 * action applied is -1 for an env whose episode is over.  rlm_env_step(h, actions != NULL) without a preceding rlm_act
 * is the "external policy" form: no generator draw is consumed.  Independent policies only. */
int rlm_act(rlm_handle h, int32_t* actions_out /* [n_envs] */);
int rlm_env_step(rlm_handle h, const int32_t* actions /* [n_envs] or NULL = the agent's own */, double* reward_out /* [n_envs] or NULL */,
                 uint8_t* terminal_out /* [n_envs] or NULL */);
int rlm_agent_update(rlm_handle h, double* delta_out /* [n_envs] TD error of each env's last transition, or NULL */);

/* measurement hooks (bench.py): CUDA-event durations of the env-tick and agent kernels, summed over launches */
int rlm_set_profiling(rlm_handle h, int32_t on);
int rlm_get_kernel_times(rlm_handle h, double* env_ms, double* agent_ms, int64_t* env_launches, int64_t* agent_launches);
/* run on a caller-provided CUDA stream (cudaStream_t as void*); 0 = the handle's own stream */
int rlm_set_stream(rlm_handle h, void* cuda_stream);

/* Real-data ingestion (host code, no GPU needed): a reference-format CSV pair -- market depth
 * (date,HH:MM:SS.mmm,AP1..5,AV1..5,BP1..5,BV1..5) and time-and-sales (date,time,price,size) -- becomes the packed message
 * stream rlm_load_ticks takes, with exactly the row filtering, print aggregation and row grouping of the reference's
 * data layer (data::basic::MarketDepth / TimeAndSales src/data/basic.cpp:20-202, Streamer::LoadUntil
 * src/data/streamer.cpp:57-81, Intraday::UpdateBookProfiles src/environment/intraday.cpp:274-313): depth rows that
 * share a timestamp or follow an invalid book state are flagged RLM_TICK_PARTIAL, ticks with more than RLM_N_TX_MAX
 * distinct print prices lead with RLM_TICK_TX_MORE messages (include/rlm_flow.h).  Call with out == NULL to size the
 * buffer: *n_msgs = messages, *n_ticks = market ticks (NextState calls) they make up. */
int rlm_ingest_csv(const char* md_path, const char* tas_path, rlm_tick_msg* out, int64_t cap, int64_t* n_msgs, int64_t* n_ticks);

/* host-side synthetic flow (same integer process as the in-kernel generator) */
int rlm_flow_generate(const rlm_flow_params* p, int64_t env_index, int64_t first_tick, int32_t n_ticks,
                      rlm_tick_msg* out);

/* ---- unit-level device entry points for the golden vectors of the reference's tests ---- */
/* Market::ToTicks / ToPrice (test/test_Market.cpp) evaluated ON THE DEVICE */
int rlm_test_to_ticks(const rlm_config* cfg, const double* px, int32_t n, int32_t* out);
int rlm_test_to_price(const rlm_config* cfg, const int32_t* ticks, int32_t n, double* out);
/* tiles() (src/rl/tiles.cpp:31-75) for n states of n_vars floats, all actions: out[n][n_actions][96] */
int rlm_test_tiles(const rlm_config* cfg, const float* vars, int32_t n, int32_t* out);
/* Order script (test/test_Order.cpp): op codes see rlm_order_op */
typedef struct rlm_order_op { int32_t op; int32_t pad; int64_t arg; } rlm_order_op; /* 0=doTransaction 1=doCancellation 2=addVolumeBehind 3=clearQueues */
typedef struct rlm_order_state { int64_t size, q_head, q_tail, executed, ret; } rlm_order_state;
int rlm_test_order(int64_t size, int64_t q_head, const rlm_order_op* ops, int32_t n_ops, rlm_order_state* out /*[n_ops]*/);
/* RollingMean<double> (test/test_Accumulators.cpp): out[i] = {mean,var} after push i */
int rlm_test_rolling_mean(int32_t window, const double* vals, int32_t n, double* out /*[n][2]*/);

#ifdef __cplusplus
}
#endif
#endif /* RLM_H */
