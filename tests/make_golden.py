"""Generate tests/golden/*.npz from the LIVE reference (dev container only).

    python tests/make_golden.py

Imports /root/reference/PyRayHF/library.py through oracle/ref_import.py (stub lmfit /
PyIRI), runs ``vertical_forward_operator`` on the cases below and stores inputs + outputs.
For every case the long-double "truth" (oracle/vfo_oracle_scalar.c variant 1, fed with the
reference's own numpy multiplier table) is stored next to the reference value, because the
float64 reference carries 3e-6 .. 5e-5 relative rounding noise in O-mode.

Files
  kat.npz        the reference's own known-answer tests (tests/test_core.py:137-152,
                 223-236, 239-276) with the values the reference produces here
  fixtures.npz   the two tutorial profiles (docs/tutorials/Example_Input_{Day,Night}.p,
                 data only) x {O,X} x n_points {1,2,50,200,2000,20000} x two frequency sets
  synthetic.npz  256 Chapman+dipole profiles (seeded subset of the 1-degree grid incl. both
                 poles and the equator) x {O,X} x n_points 200; 16 of them at 20000
  edge.npz       edge cases: B == 0, NaN in bmag, negative / zero / unsorted frequencies,
                 int altitude, non-uniform altitude grid, valley profile, near-critical sweep
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import scalar, vfo_oracle  # noqa: E402
from oracle.ref_import import load_reference_library, load_tutorial_fixture  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.simplefilter("ignore")


def run_ref(lib, f, den, bmag, bpsi, alt, mode, n):
    """Reference vh; frequencies chunked so the [F x N] temporaries stay small."""
    out = []
    for c in range(0, f.size, 174):
        out.append(lib.vertical_forward_operator(f[c:c + 174], den, bmag, bpsi, alt, mode, n))
    return np.concatenate(out)


def run_truth(f, den, bmag, bpsi, alt, mode, n):
    m = vfo_oracle.stretch_multiplier(n)
    return scalar.vertical_forward_operator(f, den, bmag, bpsi, alt, mode, n, variant=1,
                                            multiplier=m, n_threads=0, return_hc=True)


def main():
    lib = load_reference_library()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---------------- known-answer tests of the reference ----------------
    kat = {}
    X = np.array([0.02926785, 0.70981059, 0.99672596])
    Y = np.array([0.17123449, 0.16205801, 0.15757213])
    psi = np.array([60.91523271, 61.66028645, 62.02450192])
    mu, mup = lib.find_mu_mup(X, Y, psi, 'O')
    mux, mupx = lib.find_mu_mup(X, Y, psi, 'X')
    kat.update(mumup_X=X, mumup_Y=Y, mumup_psi=psi, mumup_mu_O=mu, mumup_mup_O=mup,
               mumup_mu_X=mux, mumup_mup_X=mupx,
               mumup_expected_mu=np.array([0.98626092, 0.56890941, 0.06475905]),
               mumup_expected_mup=np.array([1.01313137, 1.79819741, 19.76001084]))
    freq = np.array([1.0, 2.0, 10.0])
    alt = np.array([100, 200, 300])
    den = np.array([0, 0.5e12, 1e12])
    bm = np.array([5e-5, 5e-5, 5e-5])
    bp = np.array([60.0, 60.0, 60.0])
    kat.update(basic_freq=freq, basic_alt=alt, basic_den=den, basic_bmag=bm, basic_bpsi=bp,
               basic_vh_O=lib.vertical_forward_operator(freq, den, bm, bp, alt, 'O', 50),
               basic_vh_X=lib.vertical_forward_operator(freq, den, bm, bp, alt, 'X', 50))
    edp = np.array([5.39526842e+10, 1.77861786e+11, 6.66833260e+11])
    f3 = np.array([3.0, 3.5, 3.7])
    kat.update(model_freq=f3, model_edp=edp,
               model_expected_vh=np.array([236.22215658, 304.53151596, 334.34853791]),
               model_vh_O=lib.vertical_forward_operator(f3, edp, bm, bp, alt, 'O', 200),
               model_truth_O=run_truth(f3, edp, bm, bp, alt.astype(float), 'O', 200)[0])
    # find_vh KAT (tests/test_core.py:155-168)
    kat.update(findvh=lib.find_vh(np.array([[0.5, 0.6]]), np.array([[0.1, 0.2]]),
                                  np.array([[45.0, 45.0]]), np.array([[1.0, 1.0]]), 100.0, 'O'))
    for n in (1, 2, 10, 200, 20000):
        kat["multiplier_%d" % n] = lib.smooth_nonuniform_grid(0, 1, n, 10.)
    np.savez_compressed(os.path.join(GOLDEN, "kat.npz"), **kat)

    # ---------------- tutorial fixtures ----------------
    fx = {}
    fsets = {"a": np.arange(0.1, 17.5, 0.1), "b": np.arange(1, 16, 0.1)}
    for k, v in fsets.items():
        fx["freq_" + k] = v
    for which in ("Day", "Night"):
        d = load_tutorial_fixture(which)
        for key in ("alt", "den", "bmag", "bpsi"):
            fx["%s_%s" % (which, key)] = np.asarray(d[key], dtype=np.float64)
        for mode in "OX":
            for n in (1, 2, 50, 200, 2000, 20000):
                for fk, f in fsets.items():
                    if fk == "b" and n not in (200, 20000):
                        continue
                    tag = "%s_%s_%d_%s" % (which, mode, n, fk)
                    ref = run_ref(lib, f, d['den'], d['bmag'], d['bpsi'], d['alt'], mode, n)
                    tru, hc = run_truth(f, d['den'], d['bmag'], d['bpsi'], d['alt'], mode, n)
                    assert np.array_equal(np.isnan(ref), np.isnan(tru)), tag
                    fx["ref_" + tag] = ref
                    fx["truth_" + tag] = tru
                    fx["hc_" + tag] = hc
                    print(tag, np.isfinite(ref).sum(), "ref-truth %.2e" % np.nanmax(np.abs(ref - tru) / np.abs(tru)))
    np.savez_compressed(os.path.join(GOLDEN, "fixtures.npz"), **fx)

    # ---------------- synthetic grid subset ----------------
    sy = {}
    lat, lon = synth.grid_subset(250, seed=7)
    lat = np.concatenate([lat, [-90.0, 90.0, 0.0, 0.0, 4.5, 45.0]])
    lon = np.concatenate([lon, [0.0, 0.0, 0.0, 180.0, -150.0, 10.0]])
    alt = synth.default_alt()
    f = synth.default_freq()
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    sy.update(lat=lat, lon=lon, alt=alt, freq=f,
              input_checksum=np.array([den.sum(), bmag.sum(), bpsi.sum()]))
    for mode in "OX":
        ref = np.empty((lat.size, f.size))
        tru = np.empty_like(ref)
        for p in range(lat.size):
            ref[p] = lib.vertical_forward_operator(f, den[p], bmag[p], bpsi[p], alt, mode, 200)
            tru[p] = run_truth(f, den[p], bmag[p], bpsi[p], alt, mode, 200)[0]
        assert np.array_equal(np.isnan(ref), np.isnan(tru))
        sy["ref_%s_200" % mode] = ref
        sy["truth_%s_200" % mode] = tru
        print("synthetic", mode, 200, np.isfinite(ref).sum(),
              "ref-truth %.2e" % np.nanmax(np.abs(ref - tru) / np.abs(tru)))
        sub = np.r_[0:10, 250:256]
        ref = np.empty((sub.size, f.size))
        tru = np.empty_like(ref)
        for q, p in enumerate(sub):
            ref[q] = lib.vertical_forward_operator(f, den[p], bmag[p], bpsi[p], alt, mode, 20000)
            tru[q] = run_truth(f, den[p], bmag[p], bpsi[p], alt, mode, 20000)[0]
        sy["sub_20000"] = sub
        sy["ref_%s_20000" % mode] = ref
        sy["truth_%s_20000" % mode] = tru
        print("synthetic", mode, 20000, np.isfinite(ref).sum(),
              "ref-truth %.2e" % np.nanmax(np.abs(ref - tru) / np.abs(tru)))
    np.savez_compressed(os.path.join(GOLDEN, "synthetic.npz"), **sy)

    # ---------------- edge cases ----------------
    ed = {}
    den1, b1, p1, alt = synth.single_day_profile()
    f = synth.default_freq()

    def case(name, f, den, bmag, bpsi, alt, n=200, modes="OX"):
        ed[name + "_freq"] = np.asarray(f)
        ed[name + "_den"] = np.asarray(den)
        ed[name + "_bmag"] = np.asarray(bmag)
        ed[name + "_bpsi"] = np.asarray(bpsi)
        ed[name + "_alt"] = np.asarray(alt)
        ed[name + "_n"] = np.array(n)
        for mode in modes:
            ref = lib.vertical_forward_operator(np.asarray(f), np.asarray(den), np.asarray(bmag),
                                                np.asarray(bpsi), np.asarray(alt), mode, n)
            tru = run_truth(np.asarray(f, float), den, bmag, bpsi, np.asarray(alt, float), mode, n)[0]
            ed["%s_ref_%s" % (name, mode)] = ref
            ed["%s_truth_%s" % (name, mode)] = tru
            print(name, mode, np.isfinite(ref).sum(), "mask-eq", np.array_equal(np.isnan(ref), np.isnan(tru)))

    case("b_zero", f, den1, b1 * 0.0, p1, alt)                       # isotropic branch (lib:201-207)
    bn = b1.copy()
    bn[50] = np.nan
    case("nan_bmag", f, den1, bn, p1, alt, n=2000)                   # NaN in B below the peak
    case("odd_freq", np.array([0.0, -3.0, 3.0, 2.5, 1.0, 4.0, 0.3]), den1, b1, p1, alt)
    case("int_alt", np.array([1.0, 2.0, 10.0]), np.array([0, 0.5e12, 1e12]), np.full(3, 5e-5),
         np.full(3, 60.0), np.array([100, 200, 300]), n=50)
    # non-uniform altitude grid (geometric spacing) sampled from the same Chapman profile
    alt_nu = 80.0 + 620.0 * (np.expm1(np.linspace(0, 3, 300)) / np.expm1(3.0))
    dn, bb, pp = synth.profiles_at([30.0], [20.0], alt_nu)
    case("nonuniform_alt", f, dn[0], bb[0], pp[0], alt_nu)
    # strong E-F valley + field angle varying with height
    z = (alt - 105.0) / 6.0
    valley = den1 + 4e11 * np.exp(0.5 * (1 - z - np.exp(-z)))
    psi_var = 20.0 + 0.05 * (alt - 80.0)
    case("valley", f, valley, b1, psi_var, alt, n=2000)
    # field angle jumping by several degrees per level (forces the general sincos path)
    psi_jump = 45.0 + 20.0 * np.sin(alt / 7.0)
    case("psi_jump", f, den1, b1, psi_jump, alt, n=2000)
    # near-critical sweep (config 5): frequencies within +-0.05 MHz of the truncated-peak cutoff
    d = load_tutorial_fixture("Day")
    k = int(np.argmax(d['den']))
    fcut = np.sqrt(d['den'][k - 1]) * 8.97866275 / 1e6
    fnear = fcut + np.arange(-0.05, 0.0505, 0.001)
    case("near_crit", fnear, d['den'], d['bmag'], d['bpsi'], d['alt'], n=2000)
    case("near_crit_hi", fnear[::5], d['den'], d['bmag'], d['bpsi'], d['alt'], n=50000)
    # very short profiles
    case("two_level", np.array([1.0, 2.0, 5.0, 9.5]), np.array([1e11, 5e11, 1e12, 2e11]), np.full(4, 4e-5),
         np.array([10.0, 12.0, 14.0, 16.0]), np.array([100.0, 150.0, 250.0, 400.0]), n=500)
    case("peak_at_one", np.array([0.5, 1.0, 3.0]), np.array([1e11, 9e11, 1e10]), np.full(3, 4e-5),
         np.full(3, 30.0), np.array([100.0, 200.0, 300.0]), n=100)
    np.savez_compressed(os.path.join(GOLDEN, "edge.npz"), **ed)
    print("golden written to", GOLDEN)


if __name__ == "__main__":
    main()
