"""Reference-format CSV pair -> packed rlm_tick_msg stream.

The parser lives in the library (rl_markets_b200/csrc/rlm_ingest.cpp, C ABI rlm_ingest_csv in include/rlm.h): it
restates the reference's data layer -- row filtering and number parsing of data::basic (src/data/basic.cpp:20-202), print
aggregation of Streamer::LoadUntil (src/data/streamer.cpp:57-81), row grouping of Intraday::UpdateBookProfiles
(src/environment/intraday.cpp:274-313) -- and represents what round 1 rejected: depth rows that share a timestamp or
follow an invalid book state (RLM_TICK_PARTIAL) and ticks with more than four distinct print prices (RLM_TICK_TX_MORE).
This module is the Python face of it."""
from . import lib


def csv_pair_to_ticks(md_path, tas_path):
    """Returns (ctypes array of TickMsg, number of messages, number of market ticks they make up)."""
    return lib.ingest_csv(md_path, tas_path)


def load_day(market, md_path, tas_path):
    """Intraday::LoadData (src/environment/intraday.cpp:141-150) for every env of a STREAM-source handle: the same day for
    all of them.  Returns the number of message slots to pass to run_ticks."""
    import ctypes as C
    from . import abi
    msgs, n, _ticks = lib.ingest_csv(md_path, tas_path)
    B = market.cfg.n_envs
    wide = (abi.TickMsg * (n * B))()
    for t in range(n):
        for b in range(B):
            wide[t * B + b] = msgs[t]
    market.load_ticks(wide, n)
    return n
