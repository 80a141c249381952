"""Generates tests/golden/golden_color_v1.npz from the UNMODIFIED reference CPU engine (oracle/_ref/libdic_ref.so)
run on THREE-CHANNEL images (number_of_colors = 3): per-channel pyramid (pyramid_class.cpp:52-134) and the per-colour
loop of the evaluation (interpolation_class.cpp:712-750) with the column indexing its coefficient builders execute
(ix * (3 + channel) for bicubic / bilinear, :268-273 / :356-359).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_color.py
The domain keeps x below 0.55 * cols so that those far reads stay inside the image buffer of the reference (beyond
that the reference reads past its rows / its heap: undefined there, zero padding here)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from correlation_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_color_v1.npz")
RECT = (30, 30, 110, 150)
CENTER = (70.0, 90.0)
PYRAMID = (0, 1, 1)


def images():
    truth = (1.2, -0.7, 0.003, -0.002, 0.001, 0.004)
    chs = [synth.make_pair(200, 240, 7 + k, truth, center=(120, 100)) for k in range(3)]
    return np.stack([c[0] for c in chs], 2), np.stack([c[1] for c in chs], 2)


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref/libdic_ref.so missing: run `make -C oracle ref` first")
    und, dfm = images()
    d = {"und": und, "def": dfm, "rect": np.array(RECT, np.int32), "center": np.array(CENTER, np.float32),
         "pyramid": np.array(PYRAMID, np.int32)}
    xy = oracle.rect_points(*RECT)
    for name, interp in (("nearest", oracle.IM_NEAREST), ("bilinear", oracle.IM_BILINEAR), ("bicubic", oracle.IM_BICUBIC)):
        R = oracle.RefEngine(n_threads=3, interp=interp, pyramid=PYRAMID, colors=3)
        R.set_image("und", und)
        R.set_image("def", dfm)
        r = R.correlate(np.zeros(6, np.float32), xy, center=CENTER)
        d[f"{name}/params"] = r["params"]
        d[f"{name}/chi"] = np.float32(r["chi"])
        d[f"{name}/iterations"] = np.int32(r["iterations"])
        d[f"{name}/error_code"] = np.int32(r["error_code"])
        d[f"{name}/number_of_points"] = np.int32(r["number_of_points"])
        if interp == oracle.IM_BICUBIC:
            d["pyr_und1"] = R.pyramid_level(0, 1)
            d["pyr_def1"] = R.pyramid_level(1, 1)
            for lv in (1, 0):
                g = np.array([0.6, -0.35, 0.002, 0, 0, 0.003], np.float32)
                g[:2] *= np.float32(1.0 / (1 << lv))
                A, b, chi, err = R.evaluate(lv, g)
                d[f"bicubic/eval{lv}/A"], d[f"bicubic/eval{lv}/b"], d[f"bicubic/eval{lv}/chi"] = A, b, np.float32(chi)
        R.close()
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, len(d), "arrays")


if __name__ == "__main__":
    main()
