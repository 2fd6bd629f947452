/* rlm_flow.h -- packed tick message + synthetic Poisson order-flow generator.
 *
 * The reference (tspooner/rl_markets) has NO synthetic generator: its only
 * input is a pair of CSV files (src/data/basic.cpp:20-202).  BASELINE.json asks
 * for "synthetic Poisson order flow", so this repo defines one (SURVEY.md
 * section 8d) as a pure-integer, counter-based process that compiles
 * identically for the host (CSV writer feeding oracle/_ref, packed stream
 * feeding the oracle port and the STREAM mode of the library) and for the
 * device (GENERATOR mode: messages are produced inside the tick kernel, no HBM
 * stream at all).
 *
 * One message == one market tick == the input of one Intraday::NextState()
 * (src/environment/intraday.cpp:225-272):
 *   - the time-and-sales prints with time <= the next depth row's time,
 *     aggregated by price (data::TimeAndSalesRecord, include/data/records.h:30-38;
 *     Streamer::LoadUntil, src/data/streamer.cpp:57-81), and
 *   - that next 5-level market-depth row (data::MarketDepthRecord,
 *     include/data/records.h:20-28).
 * Prices are IEEE float because the reference parses them with stof
 * (src/data/basic.cpp:51-52,153) and only then widens to double, so a float
 * carries exactly the information the reference sees.
 */
#ifndef RLM_FLOW_H
#define RLM_FLOW_H

#include <stdint.h>

#ifndef RLM_HD
#if defined(__CUDACC__)
#define RLM_HD __host__ __device__ __forceinline__
#else
#define RLM_HD static inline
#endif
#endif

#define RLM_DEPTH 5
#define RLM_N_TX_MAX 4

/* 128-byte packed tick record (a1 in SURVEY.md section 8a). */
typedef struct rlm_tick_msg {
  float ask_px[RLM_DEPTH];    /* AP1..AP5, best first            */
  float bid_px[RLM_DEPTH];    /* BP1..BP5, best first            */
  int32_t ask_vol[RLM_DEPTH]; /* AV1..AV5                        */
  int32_t bid_vol[RLM_DEPTH]; /* BV1..BV5                        */
  float tx_px[RLM_N_TX_MAX];  /* prints aggregated by price, ascending price */
  int32_t tx_vol[RLM_N_TX_MAX];
  int32_t n_tx;               /* 0..RLM_N_TX_MAX                 */
  int32_t time_ms;            /* ms since midnight of the depth row (utilities/time.h:28-39) */
  int32_t date;               /* yyyymmdd                        */
  int32_t flags;              /* RLM_TICK_* below; 0 for one depth row = one tick (the synthetic flow) */
} rlm_tick_msg;

/* Real data does not always give one depth row per tick (rlm_ingest_csv sets these; the generator never does):
 *   RLM_TICK_PARTIAL  further depth rows of the SAME tick follow.  Intraday::UpdateBookProfiles keeps applying rows --
 *                     without stashing the book again and with the same prints -- while the next row shares the
 *                     timestamp (Streamer::WillTimeChange, src/environment/intraday.cpp:281-298)
 *                     or the book state is invalid (BookUtils::IsValidState, :300-309).  The tick completes with the
 *                     first following message that does not carry the flag.
 *   RLM_TICK_TX_MORE  the message carries no depth row, only up to RLM_N_TX_MAX further aggregated prints (ascending
 *                     price, all below the prices that follow) of the tick that the next depth row opens: a
 *                     TimeAndSalesRecord with more than RLM_N_TX_MAX distinct prices (at most RLM_TX_CAP in total). */
#define RLM_TICK_PARTIAL 1
#define RLM_TICK_TX_MORE 2
#define RLM_TX_CAP 16

/* Generator parameters.  All integer. */
typedef struct rlm_flow_params {
  uint64_t seed;        /* global seed; env b uses splitmix64(seed ^ b)           */
  int32_t mid0_tick;    /* initial best-bid tick (venue ticks, Market::ToTicks)   */
  int32_t tick_lo;      /* reflecting lower bound for the best bid tick           */
  int32_t tick_hi;      /* reflecting upper bound for the best ask tick           */
  int32_t band_tick0;   /* ticks at band_px0 (LSE AAL: 49000 at 1000.0)           */
  int32_t dt_ms;        /* ms between depth rows                                  */
  int32_t t0_ms;        /* first row is at t0_ms + dt_ms (LSE: 08:30:00.000)      */
  int32_t date;         /* yyyymmdd                                               */
  int32_t vol0;         /* initial volume per level                               */
  int32_t p_move_u12;   /* P(mid move) * 4096 (split evenly up/down)              */
  int32_t p_spread_u12; /* P(spread redraw) * 4096                                */
  int32_t spread_c1_u12;/* P(spread==1) * 4096                                    */
  int32_t spread_c2_u12;/* P(spread<=2) * 4096                                    */
  int32_t p_deep_u2;    /* prints hit the 2nd level when a 2-bit draw < this (0..4) */
  float band_px0;       /* price at band_tick0 (1000.0)                           */
  float band_ts;        /* tick size in this band (0.5)                           */
} rlm_flow_params;

/* Per-env generator state (12 ints + tick counter). */
typedef struct rlm_flow_state {
  uint32_t key0, key1;  /* philox key = per-env stream seed */
  int32_t tick;         /* index of the next message        */
  int32_t bid_tick;     /* best bid, venue ticks            */
  int32_t spread;       /* best ask - best bid, ticks       */
  int32_t ask_vol[RLM_DEPTH];
  int32_t bid_vol[RLM_DEPTH];
} rlm_flow_state;

RLM_HD uint64_t rlm_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* Philox4x32-10 (Salmon et al. 2011), counter (c0,c1,c2,c3), key (k0,k1). */
RLM_HD void rlm_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Defaults: LSE symbol AAL (src/market/market.cpp:216-227), band [1000,5000),
 * tick 0.5, start 2750.0/2750.5 (test/test_Market.cpp:26-27: ToTicks(2750.0)==52500). */
RLM_HD void rlm_flow_default_params(rlm_flow_params* p, uint64_t seed, int32_t dt_ms) {
  p->seed = seed;
  p->mid0_tick = 52500;
  p->tick_lo = 49000 + 400;
  p->tick_hi = 57000 - 400;
  p->band_tick0 = 49000;
  p->dt_ms = dt_ms;
  p->t0_ms = 8 * 3600000 + 30 * 60000;
  p->date = 20100104;
  p->vol0 = 500;
  p->p_move_u12 = 1024;   /* 0.25 */
  p->p_spread_u12 = 410;  /* ~0.1 */
  p->spread_c1_u12 = 2048;
  p->spread_c2_u12 = 3277;
  p->p_deep_u2 = 1;       /* 1/4 of prints hit the second level */
  p->band_px0 = 1000.0f;
  p->band_ts = 0.5f;
}

RLM_HD void rlm_flow_init(rlm_flow_state* s, const rlm_flow_params* p, uint64_t env_index) {
  uint64_t k = rlm_splitmix64(p->seed ^ env_index);
  s->key0 = (uint32_t)k;
  s->key1 = (uint32_t)(k >> 32);
  s->tick = 0;
  s->bid_tick = p->mid0_tick;
  s->spread = 1;
  for (int l = 0; l < RLM_DEPTH; ++l) {
    s->ask_vol[l] = p->vol0;
    s->bid_vol[l] = p->vol0;
  }
}

RLM_HD float rlm_flow_px(const rlm_flow_params* p, int32_t tick) {
  /* exact in float: small integer times 0.5 plus 1000 */
  return p->band_px0 + (float)(tick - p->band_tick0) * p->band_ts;
}

/* shift a side's level volumes by d ticks AWAY from the touch (d>0: the best
 * price moved away, deeper levels move up; d<0: new better prices appear). */
RLM_HD void rlm_flow_shift(int32_t v[RLM_DEPTH], int d, const uint32_t fresh[3]) {
  if (d > 0) {
    for (int l = 0; l < RLM_DEPTH; ++l) {
      int src = l + d;
      v[l] = (src < RLM_DEPTH) ? v[src] : (int32_t)(100u + fresh[(l + d - RLM_DEPTH) % 3] % 800u);
    }
  } else if (d < 0) {
    int m = -d;
    for (int l = RLM_DEPTH - 1; l >= 0; --l) {
      int src = l - m;
      v[l] = (src >= 0) ? v[src] : (int32_t)(100u + fresh[l % 3] % 800u);
    }
  }
}

/* Produce message number s->tick and advance the state.
 * skellam: 4096-entry int8 LUT, pois30: 4096-entry uint8 LUT, pois1p5: 256-entry uint8 LUT
 * (rlm_flow_tables.h; on the device these live in shared memory). */
/* The three Philox draws of tick s->tick (call = 0, 1, 2); independent of each other, so a warp can
 * evaluate them on three lanes. */
RLM_HD void rlm_flow_draw(const rlm_flow_state* s, uint32_t call, uint32_t out[4]) {
  rlm_philox4x32((uint32_t)s->tick, call, 0u, 0x524C4D31u, s->key0, s->key1, out);
}

RLM_HD void rlm_flow_apply(rlm_flow_state* s, const rlm_flow_params* p, const int8_t* skellam, const uint8_t* pois30,
                           const uint8_t* pois1p5, const uint32_t r0[4], const uint32_t r1[4], const uint32_t r2[4],
                           rlm_tick_msg* m);

RLM_HD void rlm_flow_next(rlm_flow_state* s, const rlm_flow_params* p, const int8_t* skellam,
                          const uint8_t* pois30, const uint8_t* pois1p5, rlm_tick_msg* m) {
  uint32_t r0[4], r1[4], r2[4];
  rlm_flow_draw(s, 0u, r0);
  rlm_flow_draw(s, 1u, r1);
  rlm_flow_draw(s, 2u, r2);
  rlm_flow_apply(s, p, skellam, pois30, pois1p5, r0, r1, r2, m);
}

RLM_HD void rlm_flow_apply(rlm_flow_state* s, const rlm_flow_params* p, const int8_t* skellam, const uint8_t* pois30,
                           const uint8_t* pois1p5, const uint32_t r0[4], const uint32_t r1[4], const uint32_t r2[4],
                           rlm_tick_msg* m) {

  /* ---- prints against the PRE-update book (prices of the previous row) ---- */
  const int32_t pa = s->bid_tick + s->spread, pb = s->bid_tick;
  int32_t agg[4] = {0, 0, 0, 0}; /* bid-1, bid, ask, ask+1 : ascending price */
  int n_prints = (s->tick == 0) ? 0 : (int)pois1p5[(r0[0] >> 24) & 0xFFu];
  for (int i = 0; i < n_prints; ++i) {
    uint32_t bits = (r0[1] >> (12 + 3 * i)) & 7u; /* bit0 side, bits1-2 depth draw */
    uint32_t u12 = (i < 2) ? ((r2[1] >> (12 * i)) & 0xFFFu) : ((r2[2] >> (12 * (i - 2))) & 0xFFFu);
    int32_t size = 1 + (int32_t)pois30[u12];
    int deep = ((int)(bits >> 1) < p->p_deep_u2) ? 1 : 0;
    if (bits & 1u) agg[2 + deep] += size; /* buy: lifts the ask   */
    else agg[1 - deep] += size;           /* sell: hits the bid   */
  }
  const int32_t agg_tick[4] = {pb - 1, pb, pa, pa + 1};
  int n_tx = 0;
  for (int i = 0; i < RLM_N_TX_MAX; ++i) { m->tx_px[i] = 0.0f; m->tx_vol[i] = 0; }
  for (int i = 0; i < 4; ++i) {
    if (agg[i] > 0) {
      m->tx_px[n_tx] = rlm_flow_px(p, agg_tick[i]);
      m->tx_vol[n_tx] = agg[i];
      ++n_tx;
    }
  }
  m->n_tx = n_tx;

  /* ---- evolve the book (not on the very first row) ---- */
  if (s->tick > 0) {
    int move = 0;
    uint32_t um = r0[0] & 0xFFFu;
    if ((int32_t)um < p->p_move_u12 / 2) move = -1;
    else if ((int32_t)um < p->p_move_u12) move = +1;
    int new_spread = s->spread;
    if ((int32_t)((r0[0] >> 12) & 0xFFFu) < p->p_spread_u12) {
      uint32_t us = r0[1] & 0xFFFu;
      new_spread = ((int32_t)us < p->spread_c1_u12) ? 1 : (((int32_t)us < p->spread_c2_u12) ? 2 : 3);
    }
    /* reflect at the band guard rails */
    if (s->bid_tick + move - (RLM_DEPTH - 1) < p->tick_lo) move = +1;
    if (s->bid_tick + move + new_spread + (RLM_DEPTH - 1) > p->tick_hi) move = -1;

    const uint32_t fresh_a[3] = {r0[2] & 0xFFFFu, r0[2] >> 16, r0[3] & 0xFFFFu};
    const uint32_t fresh_b[3] = {r0[3] >> 16, (r0[3] >> 8) & 0xFFFFu, (r0[2] >> 8) & 0xFFFFu};
    /* bid best moves by `move` (up = toward the ask = "closer": d = -move) */
    rlm_flow_shift(s->bid_vol, -move, fresh_b);
    /* ask best moves by move + (new_spread - spread) (up = away: d = +shift) */
    rlm_flow_shift(s->ask_vol, move + (new_spread - s->spread), fresh_a);
    s->bid_tick += move;
    s->spread = new_spread;

    /* depth add - cancel per level */
    for (int l = 0; l < RLM_DEPTH; ++l) {
      uint32_t w = r1[l >> 1];
      uint32_t u = (l & 1) ? ((w >> 12) & 0xFFFu) : (w & 0xFFFu);
      int32_t v = s->ask_vol[l] + (int32_t)skellam[u];
      s->ask_vol[l] = v < 1 ? 1 : v;
    }
    for (int l = 0; l < RLM_DEPTH; ++l) {
      int idx = RLM_DEPTH + l; /* draws 5..9 */
      uint32_t w = (idx < 8) ? r1[idx >> 1] : r2[0];
      uint32_t u = (idx & 1) ? ((w >> 12) & 0xFFFu) : (w & 0xFFFu);
      int32_t v = s->bid_vol[l] + (int32_t)skellam[u];
      s->bid_vol[l] = v < 1 ? 1 : v;
    }
  }

  for (int l = 0; l < RLM_DEPTH; ++l) {
    m->ask_px[l] = rlm_flow_px(p, s->bid_tick + s->spread + l);
    m->bid_px[l] = rlm_flow_px(p, s->bid_tick - l);
    m->ask_vol[l] = s->ask_vol[l];
    m->bid_vol[l] = s->bid_vol[l];
  }
  m->time_ms = p->t0_ms + (s->tick + 1) * p->dt_ms;
  m->date = p->date;
  m->flags = 0;
  s->tick += 1;
}

#endif /* RLM_FLOW_H */
