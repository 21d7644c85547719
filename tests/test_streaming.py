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
