"""The synthetic order-flow generator (include/rlm_flow.h): host renderings agree with each other,
the stream is a valid input for the reference's CSV reader, and basic invariants hold."""
import ctypes as C
import os
import subprocess
import tempfile

from rl_markets_b200 import abi, config, lib


def _cfg(seed=5, dt=250):
    return config.from_dict(config.example_dict(), flow_seed=seed, dt_ms=dt)


def test_host_entry_point_equals_standalone_writer(oracle):
    cfg = _cfg()
    a = lib.flow_generate(cfg.flow, 3, 0, 500)
    b = oracle.generate_ticks(cfg, 3, 500)  # oracle/_ref/flow_csv --packed
    assert bytes(a) == bytes(b)
    # counter-based: a later window can be generated on its own
    c = lib.flow_generate(cfg.flow, 3, 200, 100)
    assert bytes(c) == bytes(a)[200 * 128:300 * 128]
    assert bytes(lib.flow_generate(cfg.flow, 4, 0, 50)) != bytes(a)[:50 * 128]


def test_stream_invariants():
    cfg = _cfg(seed=9, dt=1)
    t = lib.flow_generate(cfg.flow, 0, 0, 20000)
    last_time = 0
    for m in t:
        assert m.time_ms > last_time and m.date == 20100104
        last_time = m.time_ms
        ap, bp = list(m.ask_px), list(m.bid_px)
        assert all(ap[i] < ap[i + 1] for i in range(4)) and all(bp[i] > bp[i + 1] for i in range(4))
        assert ap[0] > bp[0] and (ap[0] - bp[0]) in (0.5, 1.0, 1.5)
        assert all(v >= 1 for v in list(m.ask_vol) + list(m.bid_vol))
        assert 0 <= m.n_tx <= 4
        px = list(m.tx_px)[:m.n_tx]
        assert px == sorted(px) and all(v > 0 for v in list(m.tx_vol)[:m.n_tx])
        assert all((p * 2) == int(p * 2) and 1000.0 < p < 5000.0 for p in ap + bp)  # exact in float, inside the 0.5 band
    assert t[0].time_ms == 8 * 3600000 + 30 * 60000 + 1


def test_csv_rendering_round_trips(oracle):
    """flow_csv's CSV pair parses back (with stof semantics) to the packed messages."""
    cfg = _cfg(seed=2)
    n = 300
    packed = lib.flow_generate(cfg.flow, 1, 0, n)
    with tempfile.TemporaryDirectory() as d:
        md, tas = os.path.join(d, "a_md_1.csv"), os.path.join(d, "a_tas_1.csv")
        subprocess.check_call([oracle.FLOW_CSV, "--seed", "2", "--env", "1", "--ticks", str(n), "--md", md, "--tas", tas])
        rows = open(md).read().strip().splitlines()[1:]
        prints = open(tas).read().strip().splitlines()[1:]
    assert len(rows) == n
    by_time = {}
    for p in prints:
        _d, tm, px, sz = p.split(",")
        by_time.setdefault(tm, []).append((C.c_float(float(px)).value, int(sz)))
    for i, r in enumerate(rows):
        c = r.split(",")
        assert len(c) == 22
        m = packed[i]
        h, mi, s = c[1].split(":")
        ms = (int(h) * 60 + int(mi)) * 60000 + int(float(s) * 1000 + 0.5)
        assert ms == m.time_ms
        assert [C.c_float(float(x)).value for x in c[2:7]] == list(m.ask_px)
        assert [int(x) for x in c[7:12]] == list(m.ask_vol)
        assert [C.c_float(float(x)).value for x in c[12:17]] == list(m.bid_px)
        assert [int(x) for x in c[17:22]] == list(m.bid_vol)
        tx = by_time.get(c[1], [])
        assert tx == [(m.tx_px[k], m.tx_vol[k]) for k in range(m.n_tx)]


def test_csv_ingestion_reproduces_the_packed_stream(oracle):
    """rlm_ingest_csv (reference CSV pair -> packed ticks, C++ inside librlm.so, no GPU needed) inverts oracle/_ref/flow_csv."""
    cfg = _cfg(seed=17)
    n = 1500
    packed = lib.flow_generate(cfg.flow, 2, 0, n)
    with tempfile.TemporaryDirectory() as d:
        md, tas = os.path.join(d, "a_md_1.csv"), os.path.join(d, "a_tas_1.csv")
        subprocess.check_call([oracle.FLOW_CSV, "--seed", "17", "--env", "2", "--ticks", str(n), "--md", md, "--tas", tas])
        got, n_msgs, n_ticks = lib.ingest_csv(md, tas)
    assert n_msgs == n_ticks == n
    # the first message's prints are dropped by the reference (SkipUntil); the generator emits none there
    assert bytes(got) == bytes(packed)
