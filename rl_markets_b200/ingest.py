"""Reference-format CSV pair -> packed rlm_tick_msg stream (first cut).

Restates, on the host, what the reference's data layer hands to Intraday::NextState:
  * market depth rows: 22 columns date,HH:MM:SS.mmm,AP1..5,AV1..5,BP1..5,BV1..5; rows with a non-positive
    price are skipped (src/data/basic.cpp:45-70);
  * time and sales: date,time,price,size; prints with non-positive price/size are dropped
    (basic.cpp:148-162) and aggregated by 4-decimal price key into the tick whose depth-row time is
    the first one >= the print time (Streamer::LoadUntil, src/data/streamer.cpp:57-81);
  * prices are parsed with strtof exactly like the reference's stof.
Limits of the packed format (raise instead of approximating): more than RLM_N_TX_MAX distinct print prices
in one tick; several depth rows sharing a timestamp (the reference re-applies them without stashing,
SURVEY Appendix A21).
"""
import ctypes as C
import ctypes.util

from . import abi

_libc = C.CDLL(ctypes.util.find_library("c"))
_libc.strtof.restype = C.c_float
_libc.strtof.argtypes = [C.c_char_p, C.c_void_p]


def _stof(s):
    return _libc.strtof(s.encode(), None)


def _time_ms(s):  # utilities/time.h:28-39: fixed offsets HH:MM:SS.mmm
    return ((int(s[0:2]) * 60 + int(s[3:5])) * 60 + int(s[6:8])) * 1000 + int(s[9:12])


def csv_pair_to_ticks(md_path, tas_path):
    """Returns a ctypes array of TickMsg, one per accepted depth row."""
    rows = []
    with open(md_path) as f:
        next(f)  # header (csv_.skip(1), basic.cpp:25)
        for line in f:
            c = line.rstrip("\n").split(",")
            if len(c) != 22:
                continue
            ap = [_stof(x) for x in c[2:7]]
            bp = [_stof(x) for x in c[12:17]]
            if any(p <= 0.0 for p in ap + bp):
                continue
            rows.append((int(c[0]), _time_ms(c[1]), ap, [int(x) for x in c[7:12]], bp, [int(x) for x in c[17:22]]))
    prints = []
    with open(tas_path) as f:
        next(f)
        for line in f:
            c = line.rstrip("\n").split(",")
            if len(c) != 4:
                continue
            px, sz = _stof(c[2]), int(c[3])
            if px > 0.0 and sz > 0:
                prints.append((int(c[0]), _time_ms(c[1]), px, sz))
    n = len(rows)
    out = (abi.TickMsg * n)()
    j = 0
    last_t = None
    for i, (date, t, ap, av, bp, bv) in enumerate(rows):
        if last_t is not None and t == last_t:
            raise ValueError("depth rows sharing a timestamp are not representable in the packed stream (row %d)" % i)
        last_t = t
        agg = {}
        while j < len(prints) and (prints[j][0] < date or (prints[j][0] == date and prints[j][1] <= t)):
            _d, _t, px, sz = prints[j]
            key = round(float(px) * 10000)  # FloatComparator key (utilities/comparison.h:13-16)
            if key in agg:
                agg[key][1] += sz
            else:
                agg[key] = [px, sz]
            j += 1
        if len(agg) > abi.RLM_N_TX_MAX:
            raise ValueError("more than %d distinct print prices in one tick (row %d)" % (abi.RLM_N_TX_MAX, i))
        m = out[i]
        for l in range(5):
            m.ask_px[l], m.ask_vol[l], m.bid_px[l], m.bid_vol[l] = ap[l], av[l], bp[l], bv[l]
        for k, key in enumerate(sorted(agg)):
            m.tx_px[k], m.tx_vol[k] = agg[key][0], agg[key][1]
        m.n_tx = len(agg)
        m.time_ms, m.date, m.flags = t, date, 0
    return out
