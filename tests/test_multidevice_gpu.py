"""Several GPUs behind ONE jsb_run call (jsb_run_opts.devices[]): contiguous shards of the replicate
range, one host thread + stream per device, one merged result.  Replicate g uses Philox stream g wherever it runs,
so the merged result must equal the single-device one byte for byte."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = dict(Tf=60.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.02, adv_mean=0.4, adv_std=0.1,
         rule="contact", t_bottleneck=[30.0], ratio_bottleneck=[0.5], t_snapshot=[20.0])


def test_devices_list_equals_single_device(jsb):
    from jspace_b200 import _lib as L
    n = L.lib().jsb_device_count()
    devs = [0, 1] if n >= 2 else [0]
    g = jsb.Graph.lattice(40, 40, 2)
    flags = L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS
    a = jsb.run_ensemble(g, jsb.Point(**P), 37, seed=5, flags=flags)
    b = jsb.run_ensemble(g, jsb.Point(**P), 37, seed=5, flags=flags, devices=devs)
    assert a.count == b.count == 37
    assert a.summaries.tobytes() == b.summaries.tobytes()
    for i in (0, 17, 18, 19, 36):
        assert a.events(i).tobytes() == b.events(i).tobytes()
        assert a.clones(i).tobytes() == b.clones(i).tobytes()
        assert all(np.array_equal(x, y) for x, y in zip(a.lattice(i), b.lattice(i)))
        assert np.array_equal(a.list(i, 0), b.list(i, 0))
        assert a.snapshots(i).tobytes() == b.snapshots(i).tobytes()
    assert np.array_equal(a.histogram(0, 0), b.histogram(0, 0))
    a.close()
    b.close()


def test_devices_list_is_validated(jsb):
    g = jsb.Graph.lattice(10, 10, 2)
    with pytest.raises(jsb.JSpaceError):
        jsb.run_ensemble(g, jsb.Point(**P), 4, devices=[0, 0])
    with pytest.raises(jsb.JSpaceError):
        jsb.run_ensemble(g, jsb.Point(**P), 4, devices=[99])


def test_concurrent_calls_from_two_threads_share_a_graph(jsb):
    """jsb_run is re-entrant: two host threads, one graph, the same device (serialised inside the library)."""
    import threading
    g = jsb.Graph.lattice(32, 32, 2)
    ref = jsb.run_ensemble(g, jsb.Point(**P), 16, seed=9)
    out = [None, None]

    def work(k):
        out[k] = jsb.run_ensemble(g, jsb.Point(**P), 16, seed=9)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in out:
        assert r.summaries.tobytes() == ref.summaries.tobytes()
        r.close()
    ref.close()
    from jspace_b200 import _lib as L
    L.lib().jsb_trim()  # releases the cached device buffers; the next call allocates again
    jsb.run_ensemble(g, jsb.Point(**P), 4, seed=9).close()
