"""Generates tests/golden/inversion.npz from the LIVE reference (run in the dev container, where /root/reference is
mounted):  PyRayHF.library.minimize_parameters (library.py:672-825) drives its own residual_VH (library.py:594-669)
through the brute search of tests/lmfit_shim.py (scipy.optimize.brute, as lmfit does), with ``model_VH`` patched --
the way the reference's own test patches it (tests/test_core.py:345-349) -- by a Chapman profile builder, because
PyIRI's builder has no source offline.  The forward model inside is the reference's vertical_forward_operator.

    python tests/make_golden_inversion.py
"""
import os
import sys
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import  # noqa: E402
import lmfit_shim  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402

E_FO_MHZ = 3.0


def chapman_edp(nm, hm, b_bot, alt):
    nme = (E_FO_MHZ * 1e6 / synth.CP_HZ_PER_SQRT_M3) ** 2
    return synth.chapman(alt, nm, hm, b_bot) + synth.chapman(alt, nme, 110.0, 8.0)


def main():
    ref = ref_import.load_reference_library()
    lmfit_shim.install(sys.modules["lmfit"])
    ref.lmfit = sys.modules["lmfit"]
    alt = synth.default_alt()
    _, bmag, bpsi = synth.profiles_at([20.0], [0.0], alt)
    bmag, bpsi = bmag[0], bpsi[0]

    def model_vh(F2, F1, E, f_in, alt_, b_mag, b_psi, mode='O', n_points=200, bottom_type='B_bot'):
        edp = chapman_edp(float(np.squeeze(F2['Nm'])), float(np.squeeze(F2['hm'])), float(np.squeeze(F2['B_bot'])), alt_)
        return ref.vertical_forward_operator(f_in, edp, b_mag, b_psi, alt_, mode=mode, n_points=n_points), edp

    out = {"alt": alt, "bmag": bmag, "bpsi": bpsi, "e_fo_mhz": E_FO_MHZ}
    rng = np.random.default_rng(5)
    for mode in ("O", "X"):
        # observations: a profile near (but not on) a grid node, a few frequencies missing, unsorted input
        f_all = np.arange(2.0, 9.6, 0.2)
        fof2_true = 9.6 if mode == "O" else 9.0
        nm_true = (fof2_true * 1e6 / synth.CP_HZ_PER_SQRT_M3) ** 2
        vh_true = ref.vertical_forward_operator(f_all, chapman_edp(nm_true, 301.3, 47.4, alt), bmag, bpsi, alt,
                                                mode=mode, n_points=200)
        keep = np.isfinite(vh_true)
        f_obs = f_all[keep]
        vh_obs = vh_true[keep] + 0.3 * rng.standard_normal(keep.sum())
        vh_obs[3] = np.nan                                  # dropped by library.py:733
        perm = rng.permutation(f_obs.size)                  # sorted by library.py:735
        f_obs, vh_obs = f_obs[perm], vh_obs[perm]
        F2 = {"Nm": np.array([[1e12]]), "hm": np.array([[310.0]]), "B_bot": np.array([[50.0]])}
        F1 = {"P": np.array([[0.5]])}
        E = {"hm": np.array([[110.0]])}
        captured = {}
        real_minimize = lmfit_shim.minimize

        def spy(*a, **k):
            res = real_minimize(*a, **k)
            captured["res"] = res
            return res

        with mock.patch.object(ref, "model_VH", model_vh), mock.patch.object(ref.lmfit, "minimize", spy):
            vh_res, edp_res, f2_fit = ref.minimize_parameters(F2, F1, E, f_obs, vh_obs, alt, bmag, bpsi, method='brute',
                                                              percent_sigma=10., step=2., mode=mode, n_points=200)
        res = captured["res"]
        out.update({"%s_f_obs" % mode: f_obs, "%s_vh_obs" % mode: vh_obs, "%s_vh_result" % mode: vh_res,
                    "%s_edp_result" % mode: edp_res, "%s_nmf2" % mode: float(np.squeeze(f2_fit["Nm"])),
                    "%s_hmf2" % mode: float(np.squeeze(f2_fit["hm"])), "%s_b_bot" % mode: float(np.squeeze(f2_fit["B_bot"])),
                    "%s_jout" % mode: np.asarray(res.brute_Jout), "%s_grid" % mode: np.asarray(res.brute_grid)})
        print(mode, "NmF2 %.6e hmF2 %.3f B_bot %.3f  chi2 %.4f  grid %s" % (
            out["%s_nmf2" % mode], out["%s_hmf2" % mode], out["%s_b_bot" % mode], res.brute_fval,
            np.asarray(res.brute_Jout).shape))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "inversion.npz"), **out)


if __name__ == "__main__":
    main()
