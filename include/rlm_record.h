/* rlm_record.h -- per-step parity record.
 *
 * One record is emitted per completed learner step, i.e. after
 * experiment::serial::Learner::_step has run action -> performAction ->
 * newState -> HandleTransition (src/experiment/serial.cpp:53-70).  The same
 * layout is produced by
 *   - oracle/_ref/ref_driver  (the UNMODIFIED reference, compiled here),
 *   - oracle/lob_oracle       (the CPU restatement), and
 *   - the CUDA library (rlm_read_records),
 * so parity is a byte comparison of the fields marked [bit-exact] and a
 * relative-tolerance comparison (1e-5, BASELINE.json north_star) of the
 * fields marked [tol].  In practice this repo keeps the reference's operation
 * order in fp64 and compares everything bitwise.
 */
#ifndef RLM_RECORD_H
#define RLM_RECORD_H

#include <stdint.h>

#ifndef RLM_HD
#if defined(__CUDACC__)
#define RLM_HD __host__ __device__ __forceinline__
#else
#define RLM_HD static inline
#endif
#endif

#define RLM_N_STATE_MAX 13 /* environment::Variable has 13 members (include/environment/intraday.h:17-23) */

typedef struct rlm_order_rec {
  int32_t exists;        /* order_count() > 0 on this side                     */
  int32_t pad;
  double price;          /* Order::price                                        */
  int64_t q_head;        /* Order::getQueueAhead()    [bit-exact]               */
  int64_t q_tail;        /* Order::getQueueBehind()   [bit-exact]               */
  int64_t executed;      /* Order::getTotalExecutedVolume() [bit-exact]         */
} rlm_order_rec;

typedef struct rlm_step_record {
  int32_t step;          /* 0-based learner step within the episode             */
  int32_t action;        /* action taken in this step            [bit-exact]    */
  int32_t time_ms;       /* Market::time() after the step        [bit-exact]    */
  int32_t terminal;      /* Intraday::isTerminal() after the step               */
  int64_t position;      /* RiskManager::exposure()              [bit-exact]    */
  double ask_quote;      /* Base::ask_quote                      [bit-exact]    */
  double bid_quote;      /* Base::bid_quote                      [bit-exact]    */
  int32_t ask_level;     /* Intraday::ask_level                                 */
  int32_t bid_level;
  double reward;         /* env.getReward() handed to HandleTransition          */
  double pnl_step;       /* Base::pnl_step after performAction (== agg_pnl)     */
  double ep_pnl;         /* episode_stats.pnl  (cash)            [bit-exact]    */
  double ep_reward;      /* episode_stats.reward                                */
  double ep_bandh;       /* episode_stats.bandh                                 */
  /* the remaining columns of Intraday::LogProfit's profit_log row (intraday.cpp:437-451) */
  double midprice;       /* midprice(ask_book_, bid_book_) after the step       */
  double spread;         /* spread(ask_book_, bid_book_) after the step         */
  double bandh_step;     /* agg_mpm of performAction (base.cpp:285-333)         */
  rlm_order_rec ask;     /* agent ask order after the step                      */
  rlm_order_rec bid;
  int32_t ask_transactions; /* AskBook::n_transacted()           [bit-exact]    */
  int32_t bid_transactions;
  int32_t market_buys;   /* trade_stats.market_buys                             */
  int32_t market_sells;
  int32_t lo_vol_step;   /* Base::lo_vol_step                                   */
  int32_t n_state;       /* number of state variables                           */
  float state[RLM_N_STATE_MAX + 1]; /* to-state variables (Intraday::getState); backtest mode: the state
                            the action was chosen from (Backtester::_step, serial.cpp:126-128) */
  double delta;          /* TD error returned by UpdateWeights   [tol 1e-5]     */
  int32_t n_traces;      /* Traces::n_nonzero_traces after the update           */
  int32_t pad;
  uint64_t trace_hash;   /* order-independent hash of {(f, e[f], theta[f])}     */
} rlm_step_record;

/* Commutative (order-independent) accumulation of one (feature, eligibility,
 * weight) triple.  e is the float bit pattern, th the double bit pattern. */
RLM_HD uint64_t rlm_trace_mix(uint32_t f, uint32_t e_bits, uint64_t th_bits) {
  uint64_t x = ((uint64_t)f << 32) ^ (uint64_t)e_bits;
  x ^= th_bits * 0x9E3779B97F4A7C15ull;
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

#endif /* RLM_RECORD_H */
