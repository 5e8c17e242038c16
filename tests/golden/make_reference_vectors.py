"""Writes tests/golden/reference_vectors.json: the golden vectors the reference's own tests hold for the
batched linear-algebra path, transcribed from /root/reference/test/testTensor.cu with file:line citations.

Run in the build container (where /root/reference exists): every numeric literal below is checked to occur
in the cited reference test file, so the transcription cannot drift. The JSON travels to the GPU box; the
reference tree does not. Matrices are stored as flat lists in the layout the reference test uses (noted per
entry as "cm" column-major / "rm" row-major), with dims = [rows, cols, mats].
"""
import json
import os
import re
import sys

REF_TEST = "/root/reference/test/testTensor.cu"

V = {
    "_source": "GPUEngineering/GPUtils test/testTensor.cu (golden values 'from MATLAB' per the reference's comments)",
    "data_234A": {"cite": "testTensor.cu:21", "dims": [2, 3, 4], "layout": "cm",
                  "data": [1, 2, 3, 4, 5, 6, 7, 8, 9, 8, 7, 10, 5, 4, 3, 2, 1, -1, 4, 3, 4, 3, 4, 8]},
    "data_234B": {"cite": "testTensor.cu:22", "dims": [2, 3, 4], "layout": "cm",
                  "data": [7, -6, 9, 2, 1, 11, 34, -1, -4, -3, 12, 7, 9, 9, 2, 9, -9, -3, 2, 5, 4, -5, 4, 5]},
    "data_234AMB": {"cite": "testTensor.cu:24", "dims": [2, 3, 4], "layout": "cm",
                    "data": [-6, 8, -6, 2, 4, -5, -27, 9, 13, 11, -5, 3, -4, -5, 1, -7, 10, 2, 2, -2, 0, 8, 0, 3]},
    "reductions": {"cite": "testTensor.cu:488-574",
                   "dotF_A_B": 604, "normF_A": 26.153393661244042, "sumAbs_A": 112, "maxAbs_AMB": 27, "minAbs_AMB": 0},
    "addAB": {"cite": "testTensor.cu:828-848", "layout": "cm",
              "A": {"dims": [2, 3, 3], "data": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18]},
              "B": {"dims": [3, 2, 3], "data": [6, 5, 4, 3, 2, 1, 7, 6, 5, 4, 3, 2, 1, 2, 1, 5, -6, 8]},
              "C": {"dims": [2, 2, 3], "data": [41, 56, 14, 20, 158, 176, 77, 86, 60, 64, 111, 118]}},
    "transpose": {"cite": "testTensor.cu:933-945", "layout": "cm",
                  "A": {"dims": [3, 2, 2], "data": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]},
                  "At": {"dims": [2, 3, 2], "data": [1, 4, 2, 5, 3, 6, 7, 10, 8, 11, 9, 12]}},
    "least_squares_1": {"cite": "testTensor.cu:1012-1034", "layout": "cm",
                        "A": {"dims": [2, 2, 3], "data": [1, 2, 3, 4, 7, 8, 9, 10, 6, 8, -9, 20]},
                        "b": {"dims": [2, 1, 3], "data": [1, 1, -1, 2, 30, -80]}},
    "cholesky": {"cite": "testTensor.cu:1218-1276, 1377-1381", "layout": "cm (symmetric)",
                 "A": {"dims": [3, 3, 1], "data": [10.0, 2.0, 3.0, 2.0, 20.0, -1.0, 3.0, -1.0, 30.0]},
                 "L00": 3.162277660168380, "L21": -0.361403161162101, "L22": 5.382321781081287,
                 "L_rowmajor": [3.162277660168380, 0, 0, 0.632455532033676, 4.427188724235731, 0,
                                0.948683298050514, -0.361403161162101, 5.382321781081287],
                 "b": [-1.0, -3.0, 5.0],
                 "x": [-0.126805213103205, -0.128566396618528, 0.175061641423036]},
    "qr_least_squares": {"cite": "testTensor.cu:1496-1530", "layout": "rm",
                         "A": {"dims": [4, 3, 1], "data": [85.5638, -59.4001, -80.1992, 99.9464, 5.51393, 5.17935,
                                                           6.87488, -26.7536, 36.0914, -44.3857, -32.1268, 54.8915]},
                         "b": [-23.3585, -48.5744, 43.4229, -56.5081],
                         "residual_norm": 80.003169364198072},
    "svd_singular_values": {"cite": "testTensor.cu:1059-1076", "layout": "cm",
                            "B": {"dims": [8, 3, 1], "data": [1, 6, 6, 6, 6, 6, 6, 6, 2, 7, 7, 7, 7, 7, 7, 7, 3, 8, 8, 8, 8, 8, 8, 8]},
                            "S0": 32.496241123753592, "S1": 0.997152358903242},
    "svd_multiple": {"cite": "testTensor.cu:1126-1166", "layout": "cm",
                     "A": {"dims": [3, 2, 3], "data": [1, 2, 3, 4, 5, 6, 1, 1, 1, 2, 2, 2, 0, 0, 0, 0, 0, 1]},
                     "S": [9.508032000695726, 0.772869635673484, 3.872983346207417, 0, 1, 0],
                     "Vt_first4": [-0.386317703118612, -0.922365780077058, -0.922365780077058, 0.386317703118612],
                     "U": [-0.428667133548626, -0.566306918848035, -0.703946704147444,
                           0.805963908589298, 0.112382414096594, -0.581199080396110,
                           0.408248290463863, -0.816496580927726, 0.408248290463863,
                           -0.577350269189626, -0.577350269189626, -0.577350269189626,
                           0.816496580927726, -0.408248290463863, -0.408248290463863,
                           0.000000000000000, -0.707106781186548, 0.707106781186547,
                           0, 0, -1, 1, 0, 0, 0, -1, 0]},
    "svd_rank": {"cite": "testTensor.cu:1180-1193", "layout": "cm",
                 "A": {"dims": [4, 3, 3], "data": [1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 0,
                                                   1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12,
                                                   1, 2, 3, 4, 2, 4, 6, 8, 3, 6, 9, 12]},
                 "rank": [3, 2, 1]},
    "nullspace_tensor": {"cite": "testTensor.cu:1557-1586", "layout": "cm",
                         "A": {"dims": [3, 4, 5], "data": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0,
                                                           1, 2, 3, 4, 5, 6, 7, 8, 9, 7, 8, 9,
                                                           1, 2, 3, 4, 2, 4, 6, 8, 3, 6, 9, 12,
                                                           1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                                           0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]}},
    "nullspace_trivial": {"cite": "testTensor.cu:1600-1612", "layout": "rm",
                          "A": {"dims": [3, 3, 2], "data": [4, 5, 7, 4, 1, 8, 4, 5, 0, 1, 1, 1, 5, 6, 7, 9, 0, 3]}},
    "nullspace_project": {"cite": "testTensor.cu:1626-1657", "layout": "rm",
                          "A": {"dims": [3, 7, 1], "data": [1, -2, 3, 4, -1, -1, -1, 1, 2, -3, 4, -1, -1, -1, -1, 3, 5, -7, -1, -1, -1]},
                          "x": [1, 2, 3, 4, 5, 6, 7], "other": [1, -2, 5, 4, 0, 0, 0]},
    "givens_correctness": {"cite": "testTensor.cu:1715-1728", "note": "A = iota(1..60) as 10x6 cm; annihilate(0,1,2)",
                           "a00": 2.137186834969645, "a03": 44.552125559751822, "a13": -0.328797974610715},
    "tolerances": {"cite": "testTensor.cu:5-6, 1169", "PRECISION_LOW": 1e-4, "PRECISION_HIGH": 1e-10},
}


def _literals(obj):
    if isinstance(obj, dict):
        for k, v in obj.items():
            if k in ("cite", "dims", "layout", "note", "_source", "tolerances"):
                continue
            yield from _literals(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _literals(v)
    elif isinstance(obj, float) and obj != int(obj):
        yield obj


def verify_against_reference():
    text = open(REF_TEST).read()
    nums = set(float(m) for m in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", text))
    missing = [x for x in _literals(V) if x not in nums and -x not in nums]
    if missing:
        raise SystemExit(f"literals not found in {REF_TEST}: {missing[:10]}")
    print(f"verified {sum(1 for _ in _literals(V))} non-integer literals against {REF_TEST}")


if __name__ == "__main__":
    if os.path.exists(REF_TEST):
        verify_against_reference()
    else:
        print("reference tree not present: writing the JSON without re-verification", file=sys.stderr)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(V, f, indent=1)
    print("wrote", out)
