"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, its host-side pieces (model setup, tree plan, orphan pruning) agree with the oracle,
and every compute entry point fails loudly without a GPU (no fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from helpers import random_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(bnk):
    hdr = open(os.path.join(ROOT, "include", "bnkgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(bnk_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 28
    L = bnk.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(declared) == sorted(bnk.SYMBOLS)
    assert b"sm_100a" in L.bnk_version()


def test_every_declared_symbol_is_documented_for_the_reference_side():
    """INTEGRATION.md names the reference-side counterpart (or owner) of every entry point of the header, and the JNI
    shim / Java class under integration/ bind nothing the header does not declare."""
    hdr = open(os.path.join(ROOT, "include", "bnkgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bnk_[a-z_0-9]+)\s*\(", hdr))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert not sorted(s for s in declared if s not in doc)
    shim = open(os.path.join(ROOT, "integration", "jni", "bnkgpu_jni.c")).read()
    shim = re.sub(r"/\*.*?\*/", "", shim, flags=re.S)
    called = set(re.findall(r"\b(bnk_[a-z_0-9]+)\s*\(", shim))
    assert called and called <= declared, sorted(called - declared)


def test_library_has_sm100a_code_only(bnk):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "bnkit_b200", "libbnkgpu.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("name", ["LG", "JTT", "WAG", "Dayhoff", "JC", "Yang"])
def test_host_model_matches_oracle_bit_for_bit(bnk, oracle, name):
    a, b = bnk.Model(name).parts(), oracle.Model(name).parts()
    for k in ("R", "evec", "ievec", "eval", "F", "prior"):
        assert np.array_equal(a[k], b[k]), k


def test_model_variants_and_errors(bnk, oracle):
    F = np.random.default_rng(0).dirichlet(np.ones(20))
    a, b = bnk.Model("LG", F=F).parts(), oracle.Model("LG", F=F).parts()
    assert np.array_equal(a["evec"], b["evec"]) and np.array_equal(a["prior"], b["prior"])
    with pytest.raises(bnk.BnkError) as e:
        bnk.Model("NoSuchModel")
    assert e.value.code == -2
    p = bnk.Model("WAG").parts()
    m2 = bnk.Model(eigen=(20, p["F"], p["evec"], p["ievec"], p["eval"]))
    q = m2.parts(want_R=False)
    assert np.array_equal(q["evec"], p["evec"]) and np.array_equal(q["prior"], p["prior"])
    # custom model == the named one when fed the same numbers
    lg = oracle.Model("LG").parts()
    S = np.zeros((20, 20))
    iu = np.triu_indices(20, 1)
    S[iu] = (lg["R"] / lg["F"][None, :])[iu]  # any symmetric-exchangeability input works; compare product vs oracle
    c1 = bnk.Model(custom=(20, lg["F"], S, True, True)).parts()
    c2 = oracle.Model(custom=(20, lg["F"], S, True, True)).parts()
    assert np.array_equal(c1["evec"], c2["evec"]) and np.array_equal(c1["R"], c2["R"])


def test_tree_validation_and_plan(bnk):
    t = bnk.Tree([-1, 0, 0, 2, 2], [0, .1, .2, .3, .4])
    assert t.n == 5 and t.n_internal == 2 and list(t.internal_nodes()) == [0, 2]
    with pytest.raises(bnk.BnkError) as e:
        bnk.Tree([-1, 2, 0])          # parent index must be smaller than the child's
    assert e.value.code == -3
    with pytest.raises(bnk.BnkError):
        bnk.Tree([-1, 0], [0.0, -1.0])
    big = bnk.Tree(*random_tree(np.random.default_rng(3), 5000))
    assert big.n_internal == 4999


def test_prune_orphans_matches_oracle_and_fixture(bnk, oracle):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "idxtree_prune.json")))
    parent = np.array(fx["parents"], dtype=np.int32)
    t = bnk.Tree(parent)
    n = len(parent)
    rng = np.random.default_rng(5)
    masks = (rng.random((n, 64)) < 0.7).astype(np.uint8)
    for k, case in enumerate(fx["cases"]):
        masks[:, k] = 1
        masks[case["prune"], k] = 0
    out = t.prune_orphans(masks)
    for c in range(masks.shape[1]):
        want, cnt, _ = oracle.prune(parent, masks[:, c], True)
        assert np.array_equal(out[:, c], want)
    for k, case in enumerate(fx["cases"]):
        if case["same_size_with_orphan_removal"]:
            assert out[:, k].sum() == n - case["removed"]


def test_compute_entry_points_fail_loudly_without_a_gpu(bnk):
    if bnk.device_count() > 0:
        pytest.skip("a B200 is visible; the no-GPU behaviour is exercised on the build box")
    with pytest.raises(bnk.BnkError) as e:
        bnk.Workspace(bnk.Model("LG"), bnk.Tree([-1, 0, 0]), 4)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/."""
    pkg = os.path.join(ROOT, "bnkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "bnk_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_jni_shim_matches_the_header():
    """integration/jni/bnkgpu_jni.c (the binding a bnkit maintainer would add; no JDK in the image) must at least
    type-check against include/bnkgpu.h: compiled with -fsyntax-only against a stub jni.h that declares just the
    JNI calls the shim uses, and every native method of NativeInference.java must have its Java_asr_... function."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(root, "integration", "jni", "bnkgpu_jni.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "tests", "stubs"),
                        "-I", os.path.join(root, "include"), shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    java = open(os.path.join(root, "integration", "java", "asr", "NativeInference.java")).read()
    natives = re.findall(r"static native \w+ (\w+)\(", java)
    csrc = open(shim).read()
    assert len(natives) >= 12
    for name in natives:
        assert "Java_asr_NativeInference_%s(" % name in csrc, name
