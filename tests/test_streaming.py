"""Blockwise (streaming) batch MODWT with carried history -- SURVEY.md 8f row 1
(EXT/extensions/modwt/BatchStreamingMODWT.java).  CPU: the oracle restatement against the whole-signal transforms;
GPU: vectorwave_b200.BatchStreamingMODWT (vw_modwt_stream_level through the C ABI) against the oracle."""
import numpy as np
import pytest

from oracle import cref, nptwin
from oracle.wavelets import filters

REL = 1e-12


def _blocks(x, sizes):
    out, at = [], 0
    for s in sizes:
        out.append(x[:, at:at + s])
        at += s
    assert at == x.shape[1]
    return out


def test_oracle_zero_padding_stream_equals_whole_signal_transform():
    # "providing continuity across blocks and parity with whole-signal transforms" (BatchStreamingMODWT.java:14-16)
    rng = np.random.default_rng(1)
    for name, levels in (("haar", 4), ("db4", 3), ("sym8", 2)):
        h, g, _ = filters(name)
        x = rng.standard_normal((3, 700))
        so = nptwin.StreamingOracle(h, g, levels, 1)
        parts = [so.process(b) for b in _blocks(x, (100, 37, 5, 258, 300))]      # includes blocks shorter than the history
        w = np.concatenate([p[0] for p in parts], axis=2)
        v = np.concatenate([p[1] for p in parts], axis=1)
        for i in range(3):
            wo, vo = cref.decompose(x[i], h, g, levels, 1)
            assert np.array_equal(w[:, i], wo) and np.array_equal(v[i], vo)


def test_oracle_symmetric_first_block_equals_symmetric_transform_of_that_block():
    rng = np.random.default_rng(2)
    h, g, _ = filters("db4")
    x = rng.standard_normal((2, 256))
    w, v = nptwin.StreamingOracle(h, g, 3, 2).process(x)
    for i in range(2):
        wo, vo = cref.decompose(x[i], h, g, 3, 2)
        assert np.max(np.abs(w[:, i] - wo)) <= 1e-15 and np.max(np.abs(v[i] - vo)) <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name,levels", [("haar", 5), ("db4", 4), ("sym8", 3), ("coif5", 2), ("db4", 7)])
def test_streaming_blocks_match_the_oracle(mode, name, levels):
    import vectorwave_b200 as vw
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    h, g, _ = filters(name)
    rng = np.random.default_rng(levels + mode)
    x = rng.standard_normal((5, 6000))
    tol = REL * float(np.max(np.abs(x)))
    so = nptwin.StreamingOracle(h, g, levels, mode)
    with vw.BatchStreamingMODWT.Builder().wavelet(vw.get_wavelet(name)).boundary(bm).levels(levels).build() as st:
        for blk in _blocks(x, (2048, 1000, 3, 949, 2000)):          # odd lengths, a block shorter than every history
            r = st.processMultiLevel(blk)
            wo, vo = so.process(blk)
            assert float(np.max(np.abs(np.asarray(r.detailPerLevel()) - wo))) <= tol
            assert float(np.max(np.abs(np.asarray(r.finalApprox()) - vo))) <= tol
        tail = st.suggestFlushTailLength()
        assert tail == st.getMinFlushTailLength() == (len(h) - 1)
        f = st.flushMultiLevel(tail)
        wo, vo = so.flush(tail)
        assert float(np.max(np.abs(f.detailPerLevel() - wo))) <= tol and float(np.max(np.abs(f.finalApprox() - vo))) <= tol
        # flush does not advance the stream: the next block still continues the old history
        r = st.processMultiLevel(x[:, :512])
        wo, vo = so.process(x[:, :512])
        assert float(np.max(np.abs(np.asarray(r.detailPerLevel()) - wo))) <= tol


@pytest.mark.gpu
def test_streaming_single_level_periodic_state_and_errors():
    import vectorwave_b200 as vw
    from vectorwave_b200.streaming import IllegalStateException, UnsupportedOperationException
    h, g, _ = filters("db4")
    rng = np.random.default_rng(9)
    x = rng.standard_normal((4, 1024))
    # single level, ZERO_PADDING: the stream equals the whole-signal single-level transform
    st = vw.BatchStreamingMODWT.Builder().wavelet(vw.Daubechies.DB4).boundary(vw.BoundaryMode.ZERO_PADDING).build()
    a = np.concatenate([np.asarray(st.processSingleLevel(b).approx()) for b in _blocks(x, (300, 724))], axis=1)
    for i in range(4):
        v, w = cref.forward_single(x[i], h, g, 1)
        assert np.max(np.abs(a[i] - v)) <= 1e-12
    assert st.getHistoryLengthForLevel(1) == 7
    with pytest.raises(vw.IllegalArgumentException):
        st.flushSingleLevel(8)
    with pytest.raises(vw.IllegalArgumentException):
        st.getHistoryLengthForLevel(2)
    with pytest.raises(IllegalStateException):
        st.processSingleLevel(x) if False else vw.BatchStreamingMODWT.Builder().wavelet(vw.Daubechies.DB4).levels(2).build().processSingleLevel(x)
    # a batch-size change restarts the stream (BatchStreamingMODWT.java:306-320)
    r = st.processSingleLevel(x[:2, :200])
    v0, _ = cref.forward_single(x[0, :200], h, g, 1)
    assert np.max(np.abs(np.asarray(r.approx())[0] - v0)) <= 1e-12
    # PERIODIC: no state, each block on its own; flush is rejected
    sp = vw.BatchStreamingMODWT.Builder().wavelet(vw.Daubechies.DB4).levels(3).build()
    r = sp.processMultiLevel(x)
    wo, vo = cref.decompose(x[1], h, g, 3, 0)
    assert np.max(np.abs(np.asarray(r.detailPerLevel())[:, 1] - wo)) <= 1e-12
    with pytest.raises(UnsupportedOperationException):
        sp.flushMultiLevel(3)
    with pytest.raises(vw.IllegalArgumentException):
        vw.BatchStreamingMODWT.Builder().build()
    with pytest.raises(vw.IllegalArgumentException):
        st.processSingleLevel([])


# ---- core MODWTStreamingTransform (CORE/modwt/streaming/) ------------------------------------------------------------
def test_core_streaming_oracle_windows():
    o = nptwin.CoreStreamingOracle(16, 4, False)          # hop 13
    o.process(np.arange(50.0))
    assert [int(w[0]) for w in o.windows] == [0, 13, 26] and o.count == 50 - 39
    o.flush()
    assert o.windows[-1][:11].tolist() == list(np.arange(39.0, 50.0)) and not o.windows[-1][11:].any()
    m = nptwin.CoreStreamingOracle(16, 4, True)
    m.process(np.arange(40.0))
    assert [int(w[0]) for w in m.windows] == [0, 16] and m.count == 8


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_core_streaming_transform_matches_the_reference_windows(mode):
    import vectorwave_b200 as vw
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    rng = np.random.default_rng(mode)
    for name, bs in (("db4", 64), ("haar", 32), ("sym8", 256)):
        h, g, _ = filters(name)
        data = rng.standard_normal(5 * bs + 17)
        tol = 1e-12 * float(np.max(np.abs(data)))
        chunks = [data[:3], data[3:bs + 5], data[bs + 5:bs + 6], data[bs + 6:]]
        # single level: sliding windows
        got = []
        t = vw.MODWTStreamingTransform.create(vw.get_wavelet(name), bm, bs)
        t.subscribe(got.append)
        o = nptwin.CoreStreamingOracle(bs, len(h), False)
        for c in chunks:
            t.process(c)
            o.process(c)
            assert t.getBufferLevel() == o.count
        t.processSample(0.25)
        o.process([0.25])
        assert t.getStatistics().getSamplesProcessed() == data.size + 1
        assert t.getStatistics().getBlocksProcessed() == len(o.windows)
        t.close()
        o.flush()
        assert t.isClosed() and len(got) == len(o.windows)
        for r, win in zip(got, o.windows):
            v, w = nptwin.forward_single(win, h, g, mode)
            assert np.max(np.abs(r.approximationCoeffs() - v)) <= tol and np.max(np.abs(r.detailCoeffs() - w)) <= tol
        with pytest.raises(vw.InvalidStateException):
            t.process(data)
        # multi level: back-to-back windows, one result per level, approximation on the last only
        levels = 2
        got = []

        class Sub:
            done = False

            def onNext(self, r):
                got.append(r)

            def onComplete(self):
                Sub.done = True

        tm = vw.MODWTStreamingTransform.createMultiLevel(vw.get_wavelet(name), bm, bs, levels)
        tm.subscribe(Sub())
        om = nptwin.CoreStreamingOracle(bs, len(h), True)
        for c in chunks:
            tm.process(c)
            om.process(c)
        tm.flush()
        om.flush()
        assert tm.getBufferLevel() == 0 and len(got) == levels * len(om.windows)
        for k, win in enumerate(om.windows):
            w, v = nptwin.decompose(win, h, g, levels, mode)
            for level in range(1, levels + 1):
                r = got[k * levels + level - 1]
                assert np.max(np.abs(r.detailCoeffs() - w[level - 1])) <= tol
                if level == levels:
                    assert np.max(np.abs(r.approximationCoeffs() - v)) <= tol
                else:
                    assert r.approximationCoeffs().size == 0
        tm.close()
        assert Sub.done
    with pytest.raises(vw.InvalidArgumentException):
        vw.MODWTStreamingTransform.create(vw.get_wavelet("db4"), vw.BoundaryMode.PERIODIC, 4)      # bufferSize < L
    with pytest.raises(vw.InvalidArgumentException):
        vw.MODWTStreamingTransform.createMultiLevel(vw.get_wavelet("db4"), vw.BoundaryMode.PERIODIC, 64, 0)
    with pytest.raises(vw.InvalidSignalException):
        vw.MODWTStreamingTransform.create(vw.get_wavelet("db4"), vw.BoundaryMode.PERIODIC).process([])
