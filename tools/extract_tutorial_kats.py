#!/usr/bin/env python3
"""Extract the printed known answers of the reference's tutorial into a golden fixture.

Run in the build container only (needs /root/reference).  The output,
tests/golden/tutorial_lg.json, is committed; nothing reads /root/reference at test time.

Source: src/bn/ctmc/tutorial/tutorial_ML.html
  :100-103  LG stationary frequencies, 3 d.p.
  :107-129  normalised LG rate matrix, 2 d.p.
  :215-218  four unnormalised joint products of the N0/N1 example, 8 d.p.
  :225-228  the same, normalised over all 400 assignments, 3 d.p.
  :235-238  normalised with t(N0->N1) = 0.1
  :278-373  P(t) for t = 0.01, 0.02, 0.07, 0.1, 3 d.p.
"""
import json, pathlib, re

SRC = pathlib.Path("/root/reference/src/bn/ctmc/tutorial/tutorial_ML.html")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "tutorial_lg.json"
lines = SRC.read_text().splitlines()

def matrix(first, last, labelled=True):
    rows = []
    for ln in lines[first - 1:last]:
        vals = re.findall(r"-?\d+\.\d+", ln)
        if len(vals) == 20:
            rows.append([float(v) for v in vals])
    return rows

def probs(first, last):
    out = {}
    for ln in lines[first - 1:last]:
        m = re.search(r"N0=\{\\tt (\w)\}, N1=\{\\tt (\w)\}.*?= (\d\.\d+)", ln)
        if m:
            out[m.group(1) + m.group(2)] = float(m.group(3))
    return out

fix = {
    "source": "src/bn/ctmc/tutorial/tutorial_ML.html",
    "alphabet": "ARNDCQEGHILKMFPSTWYV",
    "F_3dp": matrix(100, 103)[0],
    "R_2dp": matrix(107, 129),
    "P_3dp": {"0.01": matrix(278, 301), "0.02": matrix(302, 325), "0.07": matrix(326, 349), "0.1": matrix(350, 373)},
    "example": {"tree": {"parent": [-1, 0, 0, 2], "labels": ["N0", "rno3a18", "N1", "mmus3a57"],
                         "dist": [0.0, 0.07, 0.01, 0.07]},
                "obs": {"rno3a18": "E", "mmus3a57": "D"},
                "joint_products_8dp": probs(215, 218),
                "joint_normalised_3dp": probs(225, 228),
                "joint_normalised_t01_3dp": probs(235, 238)},
}
assert len(fix["R_2dp"]) == 20 and all(len(v) == 20 for v in fix["P_3dp"].values()), "parse"
assert len(fix["example"]["joint_products_8dp"]) == 4
OUT.write_text(json.dumps(fix, indent=1))
print("wrote", OUT)
