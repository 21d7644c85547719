// Parity of the C++ host mirror (include/vw_modwt.hpp over libvwmodwt.so) against the CPU oracle (oracle/modwt_oracle.c,
// test infrastructure).  Built and run by tests/test_cpp_host.py on a GPU box; exits 0 and prints "ok" on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "vw_modwt.hpp"

extern "C" {
void vwo_forward_single(const double *x, int64_t n, const double *h, const double *g, int64_t l, int mode, double *v, double *w);
void vwo_inverse_single(const double *v, const double *w, int64_t n, const double *hr, const double *gr, int64_t l, int mode,
                        int batch_variant, double *out);
int vwo_decompose(const double *x, int64_t n, const double *h, const double *g, int64_t l, int levels, int mode, int dense,
                  double *w, double *v);
void vwo_reconstruct(const double *w, const double *v, int64_t n, const double *hr, const double *gr, int64_t l, int levels, int mode,
                     int wavelet_id, int dense, uint64_t detail_mask, int use_approx, double *out);
void vwo_batch_soa_decompose(const double *soa_x, int64_t b, int64_t n, const double *h, const double *g, int64_t l, int levels,
                             double *soa_w, double *soa_v);
void vwo_batch_soa_haar_single(const double *soa_x, int64_t b, int64_t n, double *soa_v, double *soa_w);
double vwo_swt_denoise(const double *x, int64_t n, const double *h, const double *g, int64_t l, int levels, int mode, int wavelet_id,
                       double thr, int soft, int dense, double *out);
}

using namespace vectorwave;

static int failures = 0;
static void expect_close(const std::vector<double> &a, const double *b, double tol, const char *what) {
    double m = 0;
    for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i] - b[i]));
    if (!(m <= tol)) { std::printf("FAIL %s: max diff %.3e > %.3e\n", what, m, tol); failures++; }
}

int main() {
    std::mt19937_64 rng(42);
    std::normal_distribution<double> nd;
    struct Case { Wavelet w; int oracle_id; };
    const Case cases[] = {{wavelets::haar(), 0}, {wavelets::db4(), 0}, {wavelets::db8(), 2}, {wavelets::sym8(), 4}, {wavelets::coif5(), 0}};
    const BoundaryMode modes[] = {BoundaryMode::PERIODIC, BoundaryMode::ZERO_PADDING, BoundaryMode::SYMMETRIC};
    for (const Case &c : cases) {
        const std::vector<double> h = c.w.lowPassDecomposition(), g = c.w.highPassDecomposition();
        const int l = (int)h.size();
        for (int n : {4096, 1001}) {
            std::vector<double> x(n);
            double xmax = 0;
            for (double &v : x) { v = nd(rng); xmax = std::fmax(xmax, std::fabs(v)); }
            const double tol = 1e-12 * xmax;   // north_star: 1e-12 * max|x| per level
            for (BoundaryMode m : modes) {
                const int mode = (int)m;
                // single level
                MODWTTransform t(c.w, m);
                MODWTResult r = t.forward(x);
                std::vector<double> v(n), w(n), xr(n);
                vwo_forward_single(x.data(), n, h.data(), g.data(), l, mode, v.data(), w.data());
                expect_close(r.approximationCoeffs(), v.data(), tol, "forward V");
                expect_close(r.detailCoeffs(), w.data(), tol, "forward W");
                vwo_inverse_single(v.data(), w.data(), n, h.data(), g.data(), l, mode, 0, xr.data());
                expect_close(t.inverse(MODWTResult(v, w)), xr.data(), tol, "inverse");
                // multi level
                MultiLevelMODWTTransform mt(c.w, m);
                const int levels = std::min(5, mt.getMaximumLevels(n));
                MultiLevelMODWTResult mr = mt.decompose(x, levels);
                std::vector<double> wo((size_t)levels * n), vo(n);
                vwo_decompose(x.data(), n, h.data(), g.data(), l, levels, mode, 0, wo.data(), vo.data());
                for (int j = 1; j <= levels; j++) expect_close(mr.getDetailCoeffsAtLevel(j), wo.data() + (size_t)(j - 1) * n, tol, "decompose W_j");
                expect_close(mr.getApproximationCoeffs(), vo.data(), tol, "decompose V_J");
                vwo_reconstruct(wo.data(), vo.data(), n, h.data(), g.data(), l, levels, mode, c.oracle_id, 0, (1ull << levels) - 1, 1, xr.data());
                expect_close(mt.reconstruct(mr), xr.data(), tol, "reconstruct");
                vwo_reconstruct(wo.data(), vo.data(), n, h.data(), g.data(), l, levels, mode, c.oracle_id, 0, ((1ull << levels) - 1) & ~1ull, 1, xr.data());
                expect_close(mt.reconstructFromLevel(mr, 2), xr.data(), tol, "reconstructFromLevel");
                // SWT denoise, universal soft threshold
                VectorWaveSwtAdapter swt(c.w, m);
                vwo_swt_denoise(x.data(), n, h.data(), g.data(), l, levels, mode, c.oracle_id, -1.0, 1, 0, xr.data());
                expect_close(swt.denoise(x, levels), xr.data(), tol, "swt denoise");
            }
        }
    }
    // level cap and errors (CTEST/modwt/MultiLevelMODWTTransformTest.java:271-305)
    MultiLevelMODWTTransform ht(wavelets::haar(), BoundaryMode::PERIODIC);
    if (ht.getMaximumLevels(10000) != 9) { std::printf("FAIL level cap\n"); failures++; }
    try { ht.decompose(std::vector<double>(10000, 1.0), 10); std::printf("FAIL no throw\n"); failures++; } catch (const InvalidArgumentException &) {}
    try { MODWTTransform(wavelets::db4(), BoundaryMode::CONSTANT); std::printf("FAIL CONSTANT accepted\n"); failures++; } catch (const InvalidArgumentException &) {}
    try { std::vector<double> bad(64, 1.0); bad[3] = NAN; MODWTTransform(wavelets::db4(), BoundaryMode::PERIODIC).forward(bad); std::printf("FAIL NaN accepted\n"); failures++; }
    catch (const InvalidSignalException &) {}
    // batch facade round trip (haar: exact table)
    std::vector<std::vector<double>> sig(4, std::vector<double>(512));
    for (auto &s : sig) for (double &v : s) v = nd(rng);
    auto br = BatchMODWT::multiLevelAoS(wavelets::haar(), sig, 4);
    std::vector<double> back = BatchMODWT::inverseMultiLevelAoS(wavelets::haar(), br);
    for (int i = 0; i < 4; i++) expect_close(sig[i], back.data() + (size_t)i * 512, 1e-10, "batch round trip");
    // SoA statics on the caller's layout vs the oracle's restatement of the reference SoA loop (48 lanes: column kernels)
    {
        const int b = 48, n = 640, levels = 4;
        std::vector<std::vector<double>> sg(b, std::vector<double>(n));
        for (auto &s : sg) for (double &v : s) v = nd(rng);
        std::vector<double> soa, av, wo((size_t)levels * b * n), vo((size_t)b * n);
        std::vector<std::vector<double>> dl(levels);
        BatchSIMDMODWT::convertToSoA(sg, soa);
        BatchSIMDMODWT::batchMultiLevelMODWTSoA(soa, dl, av, wavelets::db4(), b, n, levels);
        const std::vector<double> h = wavelets::db4().lowPassDecomposition(), g = wavelets::db4().highPassDecomposition();
        vwo_batch_soa_decompose(soa.data(), b, n, h.data(), g.data(), (int64_t)h.size(), levels, wo.data(), vo.data());
        for (int j = 0; j < levels; j++) expect_close(dl[j], wo.data() + (size_t)j * b * n, 1e-12 * 6.0, "SoA W_j");
        expect_close(av, vo.data(), 1e-12 * 6.0, "SoA V_J");
        std::vector<double> a1, d1;
        BatchSIMDMODWT::batchMODWTSoA(soa, a1, d1, wavelets::haar(), b, n);
        std::vector<double> v1((size_t)b * n), w1((size_t)b * n);
        vwo_batch_soa_haar_single(soa.data(), b, n, v1.data(), w1.data());
        expect_close(a1, v1.data(), 1e-12 * 6.0, "SoA haar V");
        expect_close(d1, w1.data(), 1e-12 * 6.0, "SoA haar W");
    }
    // device-resident result: levels on request, universal threshold where the coefficients are, reconstruct
    {
        const int n = 5000, levels = 4;
        std::vector<double> x(n), xr(n), wo((size_t)levels * n), vo(n);
        double xmax = 0;
        for (double &v : x) { v = nd(rng); xmax = std::fmax(xmax, std::fabs(v)); }
        const Wavelet w8 = wavelets::db8();
        const std::vector<double> h = w8.lowPassDecomposition(), g = w8.highPassDecomposition();
        for (BoundaryMode m : modes) {
            MultiLevelMODWTTransform mt(w8, m);
            ResidentResult rr(mt, x, levels);
            vwo_decompose(x.data(), n, h.data(), g.data(), (int64_t)h.size(), levels, (int)m, 0, wo.data(), vo.data());
            for (int j = 1; j <= levels; j++) expect_close(rr.getDetailCoeffsAtLevel(j), wo.data() + (size_t)(j - 1) * n, 1e-12 * xmax, "resident W_j");
            expect_close(rr.getApproximationCoeffs(), vo.data(), 1e-12 * xmax, "resident V_J");
            vwo_reconstruct(wo.data(), vo.data(), n, h.data(), g.data(), (int64_t)h.size(), levels, (int)m, 2, 0, (1ull << levels) - 1, 1, xr.data());
            expect_close(rr.reconstruct(), xr.data(), 1e-12 * xmax, "resident reconstruct");
            const double thr = rr.applyUniversalThreshold(true);
            const double tref = vwo_swt_denoise(x.data(), n, h.data(), g.data(), (int64_t)h.size(), levels, (int)m, 2, -1.0, 1, 0, xr.data());
            if (!(std::fabs(thr - tref) <= 1e-12 * tref)) { std::printf("FAIL resident universal threshold %.17g vs %.17g\n", thr, tref); failures++; }
            expect_close(rr.reconstruct(), xr.data(), 1e-12 * xmax, "resident denoise");
        }
    }
    // one long signal over three spans (all on device 0 here) driven from this one thread: equals the unsharded transform
    {
        const int world = 3, levels = 6;
        const int64_t nl = 1 << 14, n = world * nl;
        std::vector<double> x((size_t)n), xr((size_t)n), wo((size_t)levels * n), vo((size_t)n), det, app;
        double xmax = 0;
        for (double &v : x) { v = nd(rng); xmax = std::fmax(xmax, std::fabs(v)); }
        const Wavelet w4 = wavelets::db4();
        const std::vector<double> h = w4.lowPassDecomposition(), g = w4.highPassDecomposition();
        for (BoundaryMode m : {BoundaryMode::PERIODIC, BoundaryMode::ZERO_PADDING}) {
            ShardedMODWT sh(w4, m, levels, nl, std::vector<int>(world, 0));
            sh.decompose(x, det, app);
            vwo_decompose(x.data(), n, h.data(), g.data(), (int64_t)h.size(), levels, (int)m, 0, wo.data(), vo.data());
            expect_close(det, wo.data(), 1e-12 * xmax, "sharded W");
            expect_close(app, vo.data(), 1e-12 * xmax, "sharded V_J");
            vwo_reconstruct(wo.data(), vo.data(), n, h.data(), g.data(), (int64_t)h.size(), levels, (int)m, 0, 0, (1ull << levels) - 1, 1, xr.data());
            expect_close(sh.reconstruct(), xr.data(), 1e-12 * xmax, "sharded reconstruct");
        }
    }
    if (failures) { std::printf("%d failure(s)\n", failures); return 1; }
    std::printf("ok\n");
    return 0;
}
