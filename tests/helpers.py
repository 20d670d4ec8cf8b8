"""Shared helpers for the test-suite (oracle access, golden-case loading, error metric)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: 1e-5 relative (fp32), measured as max|a-b| / max|b| per tensor
RTOL_F32 = 1e-5


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = np.abs(b).max()
    if denom == 0:
        return np.abs(a).max()
    return np.abs(a - b).max() / denom


def load_dwt_cases():
    z = np.load(os.path.join(GOLDEN, "dwt_cases.npz"))
    n = int(z["ncases"])
    cases = []
    for i in range(n):
        pre = "c%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["J"] = int(case["J"])
        case["mode"] = str(case["mode"])
        case["wave"] = str(case["wave"])
        case["id"] = "%02d-%s-J%d-%s-%dx%d" % (i, case["wave"], case["J"], case["mode"],
                                               case["x"].shape[-2], case["x"].shape[-1])
        cases.append(case)
    return cases


def load_ssim_cases():
    z = np.load(os.path.join(GOLDEN, "ssim_cases.npz"))
    n = int(z["ncases"])
    cases = []
    for i in range(n):
        pre = "s%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["size_average"] = bool(case["size_average"])
        case["id"] = "%02d-sa%d" % (i, case["size_average"])
        cases.append(case)
    return cases


def case_filters(case):
    """(h_col, h_row, g_col, g_row) pairs of prepped taps as the reference's buffers hold them."""
    return ((case["h0_col"], case["h1_col"]), (case["h0_row"], case["h1_row"]),
            (case["g0_col"], case["g1_col"]), (case["g0_row"], case["g1_row"]))


def load_freq_cases():
    z = np.load(os.path.join(GOLDEN, "freq_cases.npz"))
    cases = []
    for i in range(int(z["ncases"])):
        pre = "f%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["radius"] = int(case["radius"])
        case["highpass"] = bool(case["highpass"])
        case["id"] = "%02d-%s-r%d-%dx%d" % (i, "high" if case["highpass"] else "low", case["radius"],
                                            case["x"].shape[-2], case["x"].shape[-1])
        cases.append(case)
    return cases


def load_fsd_cases():
    z = np.load(os.path.join(GOLDEN, "fsd_cases.npz"))
    cases = []
    for i in range(int(z["ncases"])):
        pre = "w%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["variant"] = str(case["variant"])
        case["cs"] = str(case["cs"])
        case["norm"] = bool(case["norm"])
        n = int(case["nbands"])
        case["y"] = [case["y%d" % b] for b in range(n)]
        case["g"] = [case["g%d" % b] for b in range(n)]
        case["id"] = "%02d-%s-%s-norm%d-%dx%d" % (i, case["variant"], case["cs"], case["norm"],
                                                  case["x"].shape[-2], case["x"].shape[-1])
        cases.append(case)
    return cases


def load_tv_cases():
    z = np.load(os.path.join(GOLDEN, "tv_cases.npz"))
    cases = []
    for i in range(int(z["ncases"])):
        pre = "t%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["weight"] = float(case["weight"])
        case["loss"] = float(case["loss"])
        case["id"] = "%02d-%s-w%g" % (i, "x".join(str(d) for d in case["x"].shape), case["weight"])
        cases.append(case)
    return cases


def load_phase_cases():
    z = np.load(os.path.join(GOLDEN, "phase_cases.npz"))
    cases = []
    for i in range(int(z["ncases"])):
        pre = "p%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["loss"] = float(case["loss"])
        case["id"] = "%02d-%s" % (i, "x".join(str(d) for d in case["x"].shape))
        cases.append(case)
    return cases


def load_dwt1d_cases():
    z = np.load(os.path.join(GOLDEN, "dwt1d_cases.npz"))
    cases = []
    for i in range(int(z["ncases"])):
        pre = "d%02d/" % i
        case = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        case["J"] = int(case["J"])
        case["mode"] = str(case["mode"])
        case["wave"] = str(case["wave"])
        case["id"] = "%02d-%s-J%d-%s-%d" % (i, case["wave"], case["J"], case["mode"], case["x"].shape[-1])
        cases.append(case)
    return cases


def load_swt_cases():
    z = np.load(os.path.join(GOLDEN, "swt_cases.npz"))
    cases = []
    pres = sorted({k.split("/")[0] for k in z.files if "/" in k})
    for pre in pres:
        case = {k[len(pre) + 1:]: z[k] for k in z.files if k.startswith(pre + "/")}
        case["mode"], case["wave"], case["dilation"] = str(case["mode"]), str(case["wave"]), int(case["dilation"])
        case["id"] = "%s-%s-%s-d%d-%dx%d" % (pre, case["wave"], case["mode"], case["dilation"], case["x"].shape[-2],
                                             case["x"].shape[-1])
        cases.append(case)
    return cases
