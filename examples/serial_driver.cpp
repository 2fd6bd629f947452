// serial_driver.cpp -- the reference's training loop (src/main.cpp:45-80 + src/experiment/serial.cpp) on the B200
// library through the class surface of include/rlm_facade.hpp: one env, one agent, N episodes.
//
//   g++ -std=c++17 -Iinclude examples/serial_driver.cpp -Lrl_markets_b200 -lrlm -Wl,-rpath,$PWD/rl_markets_b200 -o examples/serial_driver
//   examples/serial_driver [--episodes N] [--algo q_learn|sarsa|double_q_learn] [--memory-size M] [--open-ticks T] [--theta out.bin]
//
// Prints one JSON line per episode (steps, reward, pnl) and optionally dumps theta; tests/test_gpu_facade.py checks it
// against the fused rlm_run_ticks path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rlm_facade.hpp"

int main(int argc, char** argv) {
  int episodes = 2, algo = RLM_ALGO_Q_LEARN, open_ticks = 400;
  long long memory_size = 8192;
  std::string theta_out;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
    if (a == "--episodes") episodes = atoi(next().c_str());
    else if (a == "--memory-size") memory_size = atoll(next().c_str());
    else if (a == "--open-ticks") open_ticks = atoi(next().c_str());
    else if (a == "--theta") theta_out = next();
    else if (a == "--algo") { std::string v = next(); algo = v == "sarsa" ? RLM_ALGO_SARSA : (v == "double_q_learn" ? RLM_ALGO_DOUBLE_Q_LEARN : RLM_ALGO_Q_LEARN); }
  }
  try {
    rlm_config c;
    rlm::check(rlm_config_default(&c));      // config/example.yaml
    c.algorithm = algo;
    c.memory_size = memory_size;
    c.flow.seed = 41;
    c.flow.t0_ms = (int32_t)(c.close_ms - 30 * 60000 - (long long)open_ticks * c.flow.dt_ms);  // a short day: it closes after open_ticks rows
    rlm::Session session(c);
    rlm::environment::Intraday env(session);
    rlm::rl::Agent m(session);
    rlm::experiment::serial::Learner experiment(env);
    for (int episode = 1; episode <= episodes; ++episode) {   // train(), main.cpp:53-78
      env.LoadData();
      if (experiment.RunEpisode(&m))
        printf("{\"episode\": %d, \"steps\": %ld, \"reward\": %.17g, \"pnl\": %.17g, \"transactions\": %d}\n", episode, experiment.steps(),
               env.getEpisodeReward(), env.getEpisodePnL(), env.getTotalTransactions());
    }
    if (!theta_out.empty()) {
      std::vector<double> th;
      m.write_theta(th);
      FILE* f = fopen(theta_out.c_str(), "wb");
      if (!f) throw std::runtime_error("cannot open " + theta_out);
      fwrite(th.data(), 8, th.size(), f);
      fclose(f);
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "serial_driver: %s\n", e.what());
    return 1;
  }
  return 0;
}
