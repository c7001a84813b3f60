"""Loaders for golden vectors: our JSON fixtures (tests/golden) and, if ever supplied, the
reference's own downloaded files (geth CSV `input,output` rows read by src/test.c:172-241 and
rust/src/lib.rs:337-376; geth JSON read by go/blst_eip2537_test.go:18-29).  Drop the real files
into <repo>/test_vectors/ and tests/test_reference_vectors.py runs them unchanged."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "eip2537_golden.json")
REFERENCE_DIR = os.path.join(ROOT, "test_vectors")

# file name -> (ABI function, expected error code or None); names from /root/reference/build.sh:18-49
REFERENCE_FILES = {
    "g1_multiexp.csv": ("g1multiexp", None), "g2_multiexp.csv": ("g2multiexp", None), "pairing.csv": ("pairing", None),
    "g1_mul.csv": ("g1mul", None), "g2_mul.csv": ("g2mul", None),
    "g1_not_on_curve.csv": ("g1mul", 1), "g2_not_on_curve.csv": ("g2mul", 1),
    "invalid_subgroup_for_pairing.csv": ("pairing", 2),
    "blsG1MultiExp.json": ("g1multiexp", None), "blsG2MultiExp.json": ("g2multiexp", None), "blsPairing.json": ("pairing", None),
    "blsG1Mul.json": ("g1mul", None), "blsG2Mul.json": ("g2mul", None),
    "fail-blsG1MultiExp.json": ("g1multiexp", "any"), "fail-blsG2MultiExp.json": ("g2multiexp", "any"),
    "fail-blsPairing.json": ("pairing", "any"), "fail-blsG1Mul.json": ("g1mul", "any"), "fail-blsG2Mul.json": ("g2mul", "any"),
}


def load_golden():
    return json.load(open(GOLDEN))


def load_csv(path):
    rows = []
    with open(path, newline="") as f:
        for rec in csv.reader(f):
            if len(rec) >= 1 and rec[0] and rec[0].lower() != "input":
                rows.append((bytes.fromhex(rec[0].strip()), bytes.fromhex(rec[1].strip()) if len(rec) > 1 and rec[1].strip() else None))
    return rows


def load_geth_json(path):
    rows = []
    for rec in json.load(open(path)):
        exp = rec.get("Expected")
        rows.append((bytes.fromhex(rec["Input"]), bytes.fromhex(exp) if exp else None, rec.get("Name", "")))
    return rows


def reference_cases():
    """Yield (file, fn, input, expected bytes or None, expected code or 'any' or None) for supplied files."""
    if not os.path.isdir(REFERENCE_DIR):
        return
    for name, (fn, code) in REFERENCE_FILES.items():
        path = os.path.join(REFERENCE_DIR, name)
        if not os.path.exists(path):
            continue
        if name.endswith(".csv"):
            for inp, out in load_csv(path):
                yield name, fn, inp, (None if code else out), code
        else:
            for inp, out, _ in load_geth_json(path):
                yield name, fn, inp, (None if code else out), code
