"""GPU parity AT THE SIZES THE BENCHMARK NUMBERS ARE QUOTED ON ("Next round" item 1).

The full C3a / C3b / C4 workloads run on the device exactly as bench.py runs them (same generators, same seeds,
device-resident buffers through the C ABI); a sample of columns spread over all rate categories -- every node of
those columns -- is compared with the CPU oracle: joint states bit-exact, posteriors and log-likelihoods within
1e-9 RELATIVE (no absolute floor above 1e-300).  Masked variants (30 % gaps, presence mask from the device BEP
-> orphan pruning, i.e. what Prediction.getTree would hand over) take the same bar.
Reference semantics: src/bn/alg/VarElim.java:187-334; tolerance: BASELINE.json north_star.
"""
import numpy as np
import pytest

from helpers import NONE, leaves_of

pytestmark = pytest.mark.gpu
TOL = 1e-9


def sample_columns(rates, per_cat=16, seed=0):
    """column indices: `per_cat` of every distinct rate, spread over the alignment"""
    rates = np.ones(1) if rates is None else np.asarray(rates)
    rng = np.random.default_rng(seed)
    sel = []
    for r in np.unique(rates):
        idx = np.nonzero(rates == r)[0]
        sel.extend(rng.choice(idx, size=min(per_cat, len(idx)), replace=False).tolist())
    return np.array(sorted(sel), dtype=np.int64)


def max_rel_err(got, want):
    """largest |got - want| / |want| over entries with |want| >= 1e-300; NaN patterns must agree"""
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    ok = ~np.isnan(want) & (np.abs(want) >= 1e-300)
    if not ok.any():
        return 0.0
    return float((np.abs(got[ok] - want[ok]) / np.abs(want[ok])).max())


def run_and_compare(bnk, oracle, model_name, parent, dist, obs, rates, present=None, per_cat=16, query_stride=1,
                    evidence_leaves=True):
    """device-resident joint + marginal of the whole alignment, oracle on a column sample; returns stats"""
    import torch
    n, n_cols = obs.shape
    m, om = bnk.Model(model_name), oracle.Model(model_name)
    tree = bnk.Tree(parent, dist)
    dev = torch.device("cuda", 0)
    ws = bnk.Workspace(m, tree, n_cols, device=0, stream=torch.cuda.current_stream().cuda_stream)
    if evidence_leaves:
        ws.set_evidence_mode(bnk.EVIDENCE_LEAVES)
    ws.set_rates(n_cols, rates)
    d_obs = torch.from_numpy(obs).to(dev)
    d_present = None if present is None else torch.from_numpy(present).to(dev)
    d_state = torch.empty((n, n_cols), dtype=torch.uint8, device=dev)
    internal = tree.internal_nodes()
    query = None if query_stride == 1 else internal[::query_stride]
    nq = len(internal) if query is None else len(query)
    d_post = torch.empty((nq, n_cols, m.K), dtype=torch.float64, device=dev)
    d_logl = torch.empty((n_cols,), dtype=torch.float64, device=dev)
    ws.joint_dev(n_cols, d_obs, d_present, d_state)
    ws.marginal_dev(n_cols, d_obs, d_present, d_post, d_logl, query=query)
    ws.sync()
    # the one-call form bench.py times (joint pass on a second stream beside the marginal pass) must leave the very
    # same bits: same kernels, same inputs, only the scheduling differs
    d_state2, d_post2, d_logl2 = torch.full_like(d_state, 77), torch.zeros_like(d_post), torch.zeros_like(d_logl)
    ws.joint_marginal_dev(n_cols, d_obs, d_present, d_state2, d_post2, d_logl2, query=query)
    ws.sync()
    assert torch.equal(d_state2, d_state), "bnk_joint_marginal_dev: states differ from bnk_joint_dev"
    assert torch.equal(d_post2.view(torch.int64), d_post.view(torch.int64)), "bnk_joint_marginal_dev: posteriors differ"
    if present is None:
        assert torch.equal(d_logl2.view(torch.int64), d_logl.view(torch.int64)), "bnk_joint_marginal_dev: logL differs"
    else:  # per-column forests inside the wide levels: component totals are summed with atomics, in no fixed order
        assert torch.allclose(d_logl2, d_logl, rtol=1e-12, atol=0), "bnk_joint_marginal_dev: logL differs"
    del d_state2, d_post2, d_logl2
    sel = sample_columns(rates, per_cat)
    tsel = torch.from_numpy(sel).to(dev)
    st = d_state[:, tsel].cpu().numpy()
    post = d_post[:, tsel, :].cpu().numpy()
    logl = d_logl[tsel].cpu().numpy()
    sub_obs = np.ascontiguousarray(obs[:, sel])
    sub_pres = None if present is None else np.ascontiguousarray(present[:, sel])
    sub_rates = None if rates is None else np.ascontiguousarray(np.asarray(rates)[sel])
    want = oracle.fast_joint(om, parent, dist, sub_obs, sub_pres, sub_rates, nthreads=16)
    wpost, wlogl = oracle.fast_marginal(om, parent, dist, sub_obs, sub_pres, sub_rates, nthreads=16)
    wp = wpost[internal if query is None else query]
    stats = {"cols": int(len(sel)), "joint_equal": bool(np.array_equal(st, want)),
             "joint_mismatches": int((st != want).sum()),
             "post_max_rel": max_rel_err(post, wp), "logl_max_rel": max_rel_err(logl, wlogl)}
    ws.close()
    del d_post, d_state, d_obs
    torch.cuda.empty_cache()
    return stats


def check(stats):
    assert stats["joint_equal"], "joint states differ from the oracle at %d entries" % stats["joint_mismatches"]
    assert stats["post_max_rel"] <= TOL, stats
    assert stats["logl_max_rel"] <= TOL, stats


def bench_workload(name):
    """the generator bench.py uses, same seeds"""
    import bench
    spec = bench.WORKLOADS[name]
    return spec, bench.make_workload(spec, 0, spec["cols"])


def gappy(bnk, parent, dist, obs, gap_fraction, seed):
    """30 % gaps at the leaves and the presence mask the reference's pipeline would produce for them:
    device BEP (Prediction.PredictByBidirEdgeParsimony) then orphan pruning (IdxTree.getPrunedIndex)"""
    rng = np.random.default_rng(seed)
    leaf = leaves_of(parent)
    obs = obs.copy()
    obs[(rng.random(obs.shape) < gap_fraction) & leaf[:, None]] = NONE
    tree = bnk.Tree(parent, dist)
    present = tree.prune_orphans(tree.bep_presence(obs))
    return obs, present


def test_c3a_full_size_vs_oracle(bnk, oracle):
    """BASELINE configs[2], balanced: 10,000 leaves x 5,000 columns, LG + 4 rate categories; hybrid schedule
    (level kernels incl. cherry compression + walker for the narrow top); 64 columns x all 19,999 nodes."""
    spec, (model, parent, dist, rates, obs) = bench_workload("c3a")
    check(run_and_compare(bnk, oracle, spec["model"], parent, dist, obs, rates))


def test_c3b_full_size_vs_oracle(bnk, oracle):
    """BASELINE configs[2], caterpillar: 10,000 levels, walker only (hundreds of program chunks)."""
    spec, (model, parent, dist, rates, obs) = bench_workload("c3b")
    check(run_and_compare(bnk, oracle, spec["model"], parent, dist, obs, rates))


def test_c4_full_size_vs_oracle(bnk, oracle):
    """BASELINE configs[3]: 100,000-leaf random tree x 2,000 columns, WAG, one rate; posteriors of every 8th
    ancestor (the sweep still computes all of them), 24 columns x all 199,999 nodes against the oracle."""
    spec, (model, parent, dist, rates, obs) = bench_workload("c4")
    check(run_and_compare(bnk, oracle, spec["model"], parent, dist, obs, rates, per_cat=24, query_stride=8))


def test_c3a_masked_full_size_vs_oracle(bnk, oracle):
    """C3a with 30 % gaps and the BEP presence mask: the general level kernels + general walker."""
    spec, (model, parent, dist, rates, obs) = bench_workload("c3a")
    obs, present = gappy(bnk, parent, dist, obs, 0.3, 77)
    check(run_and_compare(bnk, oracle, spec["model"], parent, dist, obs, rates, present=present))


def test_c3b_masked_full_size_vs_oracle(bnk, oracle):
    """the caterpillar with 30 % gaps and the BEP presence mask: the masked deep-tree walker"""
    spec, (model, parent, dist, rates, obs) = bench_workload("c3b")
    obs, present = gappy(bnk, parent, dist, obs, 0.3, 78)
    check(run_and_compare(bnk, oracle, spec["model"], parent, dist, obs, rates, present=present, per_cat=8))


def test_host_pipeline_equals_single_shot(bnk, oracle):
    """the column-chunk pipeline of the host-buffer calls returns what one shot returns (bit for bit: same kernels,
    same association) and matches the oracle on a masked multi-category alignment with a ragged chunking"""
    from bnkit_b200 import synth
    rng = np.random.default_rng(31)
    parent, dist = synth.random_tree(300, seed=9)
    m, om = bnk.Model("LG"), oracle.Model("LG")
    n_cols = 1111
    rates = synth.gamma_category_rates(0.5, 4)[rng.integers(0, 4, n_cols)]
    obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, rates, seed=5, gap_fraction=0.2)
    leaf = leaves_of(parent)
    present = synth.path_presence_mask(parent, (obs != NONE) & leaf[:, None])
    tree = bnk.Tree(parent, dist)
    one = bnk.Workspace(m, tree, n_cols)
    one.set_pipeline(-1)
    one.set_rates(n_cols, rates)
    piped = bnk.Workspace(m, tree, n_cols)
    piped.set_pipeline(200)   # 6 chunks of 185 / 186 columns
    piped.set_rates(n_cols, rates)
    s1, s2 = one.joint(obs, present), piped.joint(obs, present)
    assert np.array_equal(s1, s2)
    assert np.array_equal(s1, oracle.fast_joint(om, parent, dist, obs, present, rates, nthreads=8))
    q = np.array([0, 5, 0, int(np.nonzero(leaf)[0][3])], dtype=np.int32)   # a duplicate and a childless node
    for query in (None, q):
        p1, l1 = one.marginal(obs, present, query=query)
        p2, l2 = piped.marginal(obs, present, query=query)
        assert np.array_equal(np.isnan(p1), np.isnan(p2))
        assert np.array_equal(np.nan_to_num(p1), np.nan_to_num(p2)) and np.array_equal(l1, l2)
    wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, rates, nthreads=8)
    assert max_rel_err(p2, wpost[q]) <= TOL and max_rel_err(l2, wlogl) <= TOL
    # both passes in one host-buffer call (bnk_joint_marginal): pipelined and single-shot, against the separate calls
    for w_ in (one, piped):
        sb, pb, lb = w_.joint_marginal(obs, present, query=q)
        assert np.array_equal(sb, s1)
        assert np.array_equal(np.isnan(pb), np.isnan(p1)) and np.array_equal(np.nan_to_num(pb), np.nan_to_num(p1))
        assert np.array_equal(lb, l1)
    ll1, t1 = one.loglik(obs, present)
    ll2, t2 = piped.loglik(obs, present)
    assert np.array_equal(ll1, ll2) and np.isclose(t1, t2, rtol=1e-13)
    assert piped.device_bytes() > 0 and piped.launch_count() > one.launch_count()
    one.close(); piped.close()
