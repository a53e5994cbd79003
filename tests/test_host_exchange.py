"""Host-side logic of the exchange primitive without a GPU: receive sides of an ExchangeGraph (default_find_rcv_ids,
src/primitives.jl:826-859) and the buffer layout of allocate_exchange (:921-947), against the oracle and the reference's
golden graph (test/primitives_tests.jl:164-188)."""
import numpy as np

from oracle import pa_oracle as o


class _HostBackend:
    """The metadata side of CUDAArray(mode='sequential'): all parts in this process."""

    def __init__(self, nparts):
        self.parts = list(range(1, nparts + 1))

    def gather_all(self, objs):
        return list(objs)


def test_find_rcv_ids_golden_and_oracle():
    import pa_b200 as pa

    b = _HostBackend(4)
    g = pa.ExchangeGraph(b, [[3, 4], [1, 3], [1, 4], [2]])
    assert g.rcv == [[2, 3], [4], [1, 2], [1, 3]]  # test/primitives_tests.jl:164-174
    rng = np.random.default_rng(0)
    for P in (1, 2, 5, 9):
        b = _HostBackend(P)
        snd = [sorted(rng.choice([q for q in range(1, P + 1) if q != p], size=rng.integers(0, P), replace=False).tolist()) if P > 1 else []
               for p in range(1, P + 1)]
        assert pa.ExchangeGraph(b, snd).rcv == o.find_rcv_ids(snd)


def test_exchange_layout_matches_oracle_exchange():
    import pa_b200 as pa

    rng = np.random.default_rng(1)
    P = 7
    b = _HostBackend(P)
    snd = [sorted(rng.choice([q for q in range(1, P + 1) if q != p], size=rng.integers(0, P - 1), replace=False).tolist()) for p in range(1, P + 1)]
    g = pa.ExchangeGraph(b, snd)
    snd_len = [[int(rng.integers(0, 9)) for _ in s] for s in snd]
    rcv_len, rcv_off, sym = pa.exchange_layout(g, snd_len)
    assert sym == max(sum(l) for l in snd_len)
    # pull every receive segment out of the senders' flat buffers with the computed offsets == the oracle's exchange
    flat = [np.arange(sum(l), dtype=np.int64) + 1000 * (p + 1) for p, l in enumerate(snd_len)]
    segs = [[flat[p][sum(l[:j]) : sum(l[: j + 1])] for j in range(len(l))] for p, l in enumerate(snd_len)]
    want = o.exchange([o.jagged_from_lists(s, np.int64) for s in segs], snd, g.rcv)
    for r in range(P):
        for i, src in enumerate(g.rcv[r]):
            got = flat[src - 1][rcv_off[r][i] : rcv_off[r][i] + rcv_len[r][i]]
            assert np.array_equal(got, want[r].segment(i))
    # an inconsistent graph is rejected like is_consistent(graph) would (src/primitives.jl:861-874)
    bad = pa.ExchangeGraph(b, snd, [[2]] + g.rcv[1:]) if 1 not in snd[1] else None
    if bad is not None:
        try:
            pa.exchange_layout(bad, snd_len)
            assert False, "expected ValueError"
        except ValueError:
            pass


def test_ghost_row_shipping_device_route_equals_host_route(monkeypatch):
    """psparse(disassembled): the entries of ghost rows reach their owners either through the metadata channel or through
    three exchange! calls; both routes must deliver the same arrays in the same (ascending sender) order.  The device
    exchange is replaced by the oracle's exchange here (no GPU); the real one is checked in tests/test_gpu_primitives.py."""
    import pa_b200 as pa
    from pa_b200 import parrays

    def fake_exchange(snd, graph):
        dtype = np.float64 if any(np.asarray(a).dtype.kind == "f" for s in snd for a in s) else np.int64
        jag = [o.jagged_from_lists([np.asarray(a, dtype=dtype) for a in s], dtype) for s in snd]
        rcv = o.exchange(jag, graph.snd, graph.rcv)
        return [[r.segment(i).copy() for i in range(len(graph.rcv[k]))] for k, r in enumerate(rcv)]

    monkeypatch.setattr(parrays, "exchange", fake_exchange)
    rng = np.random.default_rng(4)
    P = 5
    b = _HostBackend(P)
    outgoing = []
    for p in range(1, P + 1):
        dests = sorted(rng.choice([q for q in range(1, P + 1) if q != p], size=rng.integers(0, P - 1), replace=False).tolist())
        out = {}
        for q in dests:
            m = int(rng.integers(0, 6))
            out[int(q)] = (rng.integers(1, 100, m), rng.integers(1, 100, m), rng.standard_normal(m))
        outgoing.append((p, out))
    host = parrays._ship_ghost_rows(b, outgoing, "host")
    dev = parrays._ship_ghost_rows(b, outgoing, "device")
    assert len(host) == len(dev) == P
    for h, d in zip(host, dev):
        for t in range(3):
            assert h[t].dtype == d[t].dtype and np.array_equal(h[t], d[t])
