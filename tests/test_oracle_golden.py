"""Pins the CPU oracle (oracle/) to vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import bm25s_equiv, coracle
from bb25_testutil import case_scores

TOL = 1e-12  # oracle (libm) vs reference (numpy SIMD exp/log): a few ulp


def test_survey_known_answers(golden_pf):
    g = golden_pf
    # SURVEY 8c table, independent of the npz
    np.testing.assert_allclose(
        g["g1"], [0.00300322718859175, 0.01032184655396619, 0.01712686496535638,
                  0.03934663032138969, 0.25028854884930135], rtol=0, atol=1e-15)
    np.testing.assert_allclose(g["g5"], [0.02302631578947368], atol=1e-15)
    o = coracle.score_to_probability(1.5, 1.0, 0.01, [.5, 1, 1.5, 2, 3], [1, 2, 3, 5, 8], [.3, .5, .8, 1, 1.5])
    np.testing.assert_allclose(o, g["g1"], rtol=0, atol=TOL)
    o = coracle.score_to_probability(1.0, 0.0, None, [1.0464478, 0.56150854, 1.1230172], [5, 3, 7], [.5, .5, .5])
    np.testing.assert_allclose(o, g["g2"], rtol=0, atol=TOL)
    o = coracle.score_to_probability(2.020913298430858, 0.2106551229953766, .04501,
                                     np.array([3.1415927, .001, 12.5, 25, 60], dtype=np.float32),
                                     [1, 2, 3, 4, 4], [.2, .7, 1, 1.3, 3])
    np.testing.assert_allclose(o, g["g6"], rtol=0, atol=TOL)
    assert coracle.wand_upper_bound(1.5, 2.0, .01, [5.0])[0] == pytest.approx(g["g4"][0], abs=TOL)
    assert coracle.wand_upper_bound(1.5, 2.0, None, [5.0])[0] == pytest.approx(g["g4"][1], abs=TOL)
    assert coracle.lib().orc_posterior(0.7, 0.5, 1, 0.01) == pytest.approx(g["g5"][0], abs=TOL)
    assert coracle.log_odds_conjunction([.85, .7, .6]) == pytest.approx(g["g7"][0], abs=TOL)
    np.testing.assert_allclose(coracle.log_odds_conjunction(g["g8_in"], weights=[.6, .4]), g["g8a"], atol=TOL)
    np.testing.assert_allclose(coracle.log_odds_conjunction(g["g8_in"], alpha=.5, weights=[.6, .4]), g["g8b"], atol=TOL)
    for i, gt in enumerate(("relu", "swish", "gelu", "softplus")):
        assert coracle.log_odds_conjunction(g["g8_in"], gating=gt)[0] == pytest.approx(g["g10"][i], abs=TOL)
    assert coracle.lib().orc_logit(1e-10) == pytest.approx(-23.025850929840455, abs=1e-12)


def test_probability_sweeps(golden_pf):
    g = golden_pf
    s, tf, r = g["sweep_score"], g["sweep_tf"], g["sweep_ratio"]
    for i, (a, b, br) in enumerate(g["sweep_params"]):
        br = None if br < 0 else float(br)
        np.testing.assert_allclose(coracle.score_to_probability(a, b, br, s, tf, r),
                                   g[f"sweep_prob_{i}"], rtol=0, atol=TOL)
        np.testing.assert_allclose(coracle.score_to_probability(a, b, br, s, tf, r, prior_mode=1),
                                   g[f"sweep_priorfree_{i}"], rtol=0, atol=TOL)
        np.testing.assert_allclose(coracle.wand_upper_bound(a, b, br, s.astype(np.float64)),
                                   g[f"sweep_wand_{i}"], rtol=0, atol=TOL)
    L = coracle.lib()
    np.testing.assert_allclose([L.orc_tf_prior(float(x)) for x in tf], g["sweep_tf_prior"], atol=1e-15)
    np.testing.assert_allclose([L.orc_norm_prior(float(x)) for x in r], g["sweep_norm_prior"], atol=1e-15)
    np.testing.assert_allclose([L.orc_composite_prior(float(a), float(b)) for a, b in zip(tf, r)],
                               g["sweep_composite"], atol=1e-15)
    np.testing.assert_allclose([L.orc_posterior(float(a), float(b), 0, 0.0) for a, b in zip(g["post_l"], g["post_p"])],
                               g["post_nobr"], atol=1e-15)
    np.testing.assert_allclose([L.orc_posterior(float(a), float(b), 1, 0.02) for a, b in zip(g["post_l"], g["post_p"])],
                               g["post_br"], atol=1e-15)
    np.testing.assert_allclose([L.orc_sigmoid(float(x)) for x in g["sig_x"]], g["sig_y"], atol=1e-15)
    np.testing.assert_allclose([L.orc_logit(float(x)) for x in g["logit_p"]], g["logit_y"], atol=1e-12)
    np.testing.assert_allclose(coracle.cosine_to_probability(g["cos_x"]), g["cos_y"], atol=0)


@pytest.mark.parametrize("nsig", [1, 2, 3, 5, 9])
def test_log_odds_conjunction_sweeps(golden_pf, nsig):
    g = golden_pf
    P, w = g[f"loc_in_{nsig}"], g[f"loc_w_{nsig}"]
    chk = lambda got, key: np.testing.assert_allclose(got, g[key], rtol=0, atol=TOL)
    chk(coracle.log_odds_conjunction(P), f"loc_unw_{nsig}")
    chk(coracle.log_odds_conjunction(P, alpha=0.0), f"loc_unw_a0_{nsig}")
    chk(coracle.log_odds_conjunction(P, alpha="auto"), f"loc_unw_auto_{nsig}")
    chk(coracle.log_odds_conjunction(P, weights=w), f"loc_w_{nsig}_none")
    chk(coracle.log_odds_conjunction(P, alpha=0.5, weights=w), f"loc_w_{nsig}_a05")
    for gt in ("relu", "swish", "gelu", "softplus"):
        chk(coracle.log_odds_conjunction(P, alpha=0.5, weights=w, gating=gt), f"loc_{gt}_{nsig}")
        chk(coracle.log_odds_conjunction(P, gating=gt, gating_beta=2.0), f"loc_{gt}_b2_{nsig}")
    chk(coracle.log_odds_conjunction(P, weights=w, max_logit=3.0), f"loc_clip_{nsig}")


def test_scorer_cases_against_reference(golden_scorer):
    """C oracle pipeline == reference scorer.py output on the same CSC."""
    arrays, metas = golden_scorer
    for m in metas:
        p = m["prefix"]
        sc = case_scores(arrays, m)
        params = coracle.make_params(m["alpha"], m["beta"], m["base_rate"])
        qt, qo = arrays[p + "q_terms"], arrays[p + "q_off"]
        for k in m["ks"]:
            if f"{p}ids_k{k}" not in arrays:
                continue
            ids, scs, probs, _ = coracle.retrieve_batch(sc, params, qt, qo, k, n_threads=2)
            np.testing.assert_array_equal(ids, arrays[f"{p}ids_k{k}"], err_msg=m["name"])
            np.testing.assert_array_equal(scs, arrays[f"{p}scores_k{k}"])
            np.testing.assert_allclose(probs, arrays[f"{p}probs_k{k}"], rtol=0, atol=TOL)
        for i in range(m["n_dense"]):
            q = qt[qo[i]:qo[i + 1]]
            np.testing.assert_array_equal(coracle.get_scores(sc, q), arrays[p + "dense_scores"][i])
            np.testing.assert_array_equal(bm25s_equiv.get_scores_ids(sc, q), arrays[p + "dense_scores"][i])
            np.testing.assert_allclose(coracle.get_probabilities(sc, params, q),
                                       arrays[p + "dense_probs"][i], rtol=0, atol=TOL)
            np.testing.assert_array_equal(coracle.match_counts(sc, q), bm25s_equiv.match_counts(sc, q))


def test_blockmax_and_merge(golden_mf):
    arrays, _ = golden_mf
    sm = arrays["bmw_matrix"]
    for bs in (1, 7, 128, 1000, 4096):
        np.testing.assert_array_equal(coracle.blockmax_dense(sm, bs), arrays[f"bmw_bs{bs}"])
    bm = coracle.blockmax_dense(sm, 128)
    np.testing.assert_allclose(coracle.wand_upper_bound(1.3, 1.7, 0.03, bm).reshape(bm.shape),
                               arrays["bmw_bayes_bs128"], rtol=0, atol=TOL)
    # merge of shard-local lists == top-k of the union
    rng = np.random.default_rng(0)
    S, Q, k, N = 4, 5, 16, 400
    scores = np.round(rng.uniform(0, 2, (Q, N)), 1).astype(np.float32)  # many ties
    ids = np.empty((S, Q, k), np.int64); sc = np.empty((S, Q, k), np.float32); pr = np.empty((S, Q, k))
    for s in range(S):
        lo, hi = s * N // S, (s + 1) * N // S
        for q in range(Q):
            i, v = coracle.topk_f32(scores[q, lo:hi], k)
            ids[s, q], sc[s, q], pr[s, q] = i + lo, v, v * 0.5
    oi, os_, op = coracle.merge_topk(ids, sc, pr)
    for q in range(Q):
        i, v = coracle.topk_f32(scores[q], k)
        np.testing.assert_array_equal(oi[q], i)
        np.testing.assert_array_equal(os_[q], v)
        np.testing.assert_array_equal(op[q], v * 0.5)
