// flow_csv -- TEST INFRASTRUCTURE ONLY.
//
// Renders the synthetic order flow of include/rlm_flow.h for ONE env index as
//   (a) the reference's CSV pair: market depth = header + 22 columns
//       date,HH:MM:SS.mmm,AP1..AP5,AV1..AV5,BP1..BP5,BV1..BV5
//       (/root/reference/include/data/basic.h:17-24, src/data/basic.cpp:45-70) and
//       time-and-sales = header + date,time,price,size (basic.h:49-52, basic.cpp:148-162);
//   (b) optionally the packed rlm_tick_msg stream (binary, 128 B per tick).
// Every print of tick t carries the timestamp of depth row t so that
// Streamer::LoadUntil (src/data/streamer.cpp:57-81) attaches it to exactly
// that NextState() call.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rlm_flow.h"
#include "rlm_flow_tables.h"

static void fmt_time(int32_t ms, char* out) {
  int mil = ms % 1000; ms /= 1000;
  int sec = ms % 60; ms /= 60;
  int min = ms % 60; ms /= 60;
  sprintf(out, "%02d:%02d:%02d.%03d", ms, min, sec, mil);
}

int main(int argc, char** argv) {
  uint64_t seed = 1;
  uint64_t env = 0;
  long n_ticks = 1000;
  int dt_ms = 250;
  long t0_ms = -1;
  std::string md, tas, packed;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a == "--seed") seed = strtoull(next(), 0, 10);
    else if (a == "--env") env = strtoull(next(), 0, 10);
    else if (a == "--ticks") n_ticks = atol(next());
    else if (a == "--dt-ms") dt_ms = atoi(next());
    else if (a == "--t0-ms") t0_ms = atol(next());
    else if (a == "--md") md = next();
    else if (a == "--tas") tas = next();
    else if (a == "--packed") packed = next();
    else { fprintf(stderr, "usage: flow_csv --seed S --env B --ticks N [--dt-ms D] [--t0-ms T] [--md f] [--tas f] [--packed f]\n"); return 2; }
  }
  rlm_flow_params p;
  rlm_flow_default_params(&p, seed, dt_ms);
  if (t0_ms >= 0) p.t0_ms = (int32_t)t0_ms;
  rlm_flow_state s;
  rlm_flow_init(&s, &p, env);

  FILE* fm = md.empty() ? nullptr : fopen(md.c_str(), "w");
  FILE* ft = tas.empty() ? nullptr : fopen(tas.c_str(), "w");
  FILE* fp = packed.empty() ? nullptr : fopen(packed.c_str(), "wb");
  if ((!md.empty() && !fm) || (!tas.empty() && !ft) || (!packed.empty() && !fp)) { perror("flow_csv: open"); return 1; }
  if (fm) fprintf(fm, "date,time,ap1,ap2,ap3,ap4,ap5,av1,av2,av3,av4,av5,bp1,bp2,bp3,bp4,bp5,bv1,bv2,bv3,bv4,bv5\n");
  if (ft) fprintf(ft, "date,time,price,size\n");

  char tbuf[32];
  rlm_tick_msg m;
  for (long t = 0; t < n_ticks; ++t) {
    rlm_flow_next(&s, &p, rlm_flow_skellam20_lut, rlm_flow_pois30_lut, rlm_flow_pois1p5_lut, &m);
    fmt_time(m.time_ms, tbuf);
    if (fp) fwrite(&m, sizeof(m), 1, fp);
    if (ft)
      for (int i = 0; i < m.n_tx; ++i) fprintf(ft, "%d,%s,%.4f,%d\n", m.date, tbuf, (double)m.tx_px[i], m.tx_vol[i]);
    if (fm) {
      fprintf(fm, "%d,%s", m.date, tbuf);
      for (int l = 0; l < RLM_DEPTH; ++l) fprintf(fm, ",%.4f", (double)m.ask_px[l]);
      for (int l = 0; l < RLM_DEPTH; ++l) fprintf(fm, ",%d", m.ask_vol[l]);
      for (int l = 0; l < RLM_DEPTH; ++l) fprintf(fm, ",%.4f", (double)m.bid_px[l]);
      for (int l = 0; l < RLM_DEPTH; ++l) fprintf(fm, ",%d", m.bid_vol[l]);
      fprintf(fm, "\n");
    }
  }
  if (fm) fclose(fm);
  if (ft) fclose(ft);
  if (fp) fclose(fp);
  return 0;
}
