#!/usr/bin/env python3
"""Turn a clean synthetic CSV pair (oracle/_ref/flow_csv) into one with the irregularities of real data that the packed
stream must represent: depth rows sharing a timestamp (Appendix A21), intervals with more than four
distinct print prices, depth rows with a zero price (dropped by data::basic::MarketDepth::_ParseRow), rows of the wrong
width, and a crossed book (an invalid state: the reference swallows the following row into the same tick).
Test infrastructure: used by tools/make_golden.py and tests/test_ingest.py."""
import random


def make_messy(md_in, tas_in, md_out, tas_out, seed=1, start=120, features=("dup", "zero", "short", "cross", "burst")):
    rnd = random.Random(seed)
    md = open(md_in).read().split("\n")
    tas = open(tas_in).read().split("\n")
    head, rows = md[0], [r for r in md[1:] if r]
    out = [head]
    extra_prints = []
    for i, r in enumerate(rows):
        c = r.split(",")
        out.append(r)
        if i < start:
            continue
        u = rnd.random()
        if u < 0.06 and "dup" in features:      # a second (and sometimes third) row with the same timestamp, other volumes
            for _ in range(1 + (rnd.random() < 0.3)):
                d = list(c)
                for k in list(range(7, 12)) + list(range(17, 22)):
                    d[k] = str(max(1, int(d[k]) + rnd.randint(-40, 40)))
                out.append(",".join(d))
        elif 0.06 <= u < 0.09 and "zero" in features:    # a row with a zero price: dropped by the reader
            d = list(c)
            d[2 + rnd.randint(0, 4)] = "0.0000"
            d[1] = c[1][:-1] + "1"  # (its own timestamp, 1 ms later)
            out.append(",".join(d))
        elif 0.09 <= u < 0.11 and "short" in features:    # a truncated line
            out.append(",".join(c[:9]))
        elif 0.11 <= u < 0.13 and "cross" in features:    # a crossed book 1 ms later: best ask below best bid -> IsValidState false -> the next row joins the tick
            d = list(c)
            bb = float(c[12])
            for l in range(5):
                d[2 + l] = "%.4f" % (bb - 1.0 + 0.5 * l)
            d[1] = c[1][:-1] + "2"
            out.append(",".join(d))
        if rnd.random() < 0.05 and "burst" in features:  # a burst of prints at 5..9 distinct prices inside this row's interval
            ba, bb = float(c[2]), float(c[12])
            n = rnd.randint(5, 9)
            pxs = rnd.sample([bb - 0.5 * k for k in range(0, 6)] + [ba + 0.5 * k for k in range(0, 6)], n)
            for p in pxs:
                extra_prints.append((c[0], c[1], "%.4f" % p, str(rnd.randint(1, 80))))
    with open(md_out, "w") as f:
        f.write("\n".join(out) + "\n")
    prints = [tuple(p.split(",")) for p in tas[1:] if p] + extra_prints
    prints.sort(key=lambda p: (p[0], p[1]))  # stable: same-time prints keep their order
    with open(tas_out, "w") as f:
        f.write(tas[0] + "\n" + "\n".join(",".join(p) for p in prints) + "\n")
