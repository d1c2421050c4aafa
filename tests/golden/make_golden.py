#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference imported from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

`matplotlib` and `pykalman` (imported by the reference's test models but absent from this image)
are stubbed in sys.modules; nothing on the recorded path touches them.

Two numpy CPU-dispatch variants are recorded, because the reference's bits depend on it
(DESIGN.md "parity pinning"): scipy.special.logsumexp calls np.log1p, which numpy serves from
Intel SVML on AVX-512 hosts and from glibc's log1pf elsewhere.
    variant "default" : numpy as it dispatches on this host (AVX512_SKX -> SVML log1p)
    variant "avx2"    : NPY_DISABLE_CPU_FEATURES = all AVX-512 groups (glibc log1pf; exp/log loops
                        are the same algorithm on AVX2 and AVX-512)
The script re-executes itself once per variant and merges the results.
"""
import json
import os
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
AVX512 = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "pykalman"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import aesmc  # noqa: F401  (the reference package)
    import aesmc.inference, aesmc.math, aesmc.state, aesmc.statistics  # noqa: F401,E401
    from test.models import lgssm
    return aesmc, lgssm


def step_cases():
    """(name, B, K, generator) for sample_ancestral_index replay."""
    import numpy as np

    def normal(scale, shift=-1.4):
        def g(rng, B, K):
            return (rng.standard_normal((B, K)) * scale + shift).astype(np.float32)
        return g

    def equal(rng, B, K):
        return np.full((B, K), -0.5, np.float32)

    def with_neg_inf(rng, B, K):
        x = (rng.standard_normal((B, K)) - 1.4).astype(np.float32)
        x[:, ::3] = -np.inf
        return x

    def tied_max(rng, B, K):
        x = (rng.standard_normal((B, K)) - 1.4).astype(np.float32)
        x[:, 1] = x.max(axis=1)
        x[:, -1] = x.max(axis=1)
        return x

    def peaked(rng, B, K):
        x = (rng.standard_normal((B, K)) * 0.1 - 60.0).astype(np.float32)
        x[:, K // 2] = 3.0
        return x

    cases = [
        ("k1", 3, 1, normal(1.0)), ("k2", 3, 2, normal(1.0)), ("k3", 5, 3, normal(1.0)),
        ("k7", 4, 7, normal(1.0)), ("k100", 4, 100, normal(1.0)), ("k129", 4, 129, normal(1.0)),
        ("k1000", 8, 1000, normal(1.0)), ("k1000_heavy", 8, 1000, normal(5.0)),
        ("k4096", 6, 4096, normal(1.0)), ("k4096_heavy", 6, 4096, normal(5.0)),
        ("k4097", 2, 4097, normal(1.0)), ("k8192", 2, 8192, normal(1.0)),
        ("k20000", 1, 20000, normal(1.0)), ("k65536", 1, 65536, normal(1.0)),
        ("equal_k64", 2, 64, equal), ("equal_k4096", 1, 4096, equal),
        ("neginf_k512", 3, 512, with_neg_inf), ("tiedmax_k300", 3, 300, tied_max),
        ("peaked_k2048", 2, 2048, peaked),
    ]
    return cases


def record(variant):
    import numpy as np
    import scipy
    import scipy.special
    import torch
    aesmc, lgssm = _import_reference()
    out = {}
    meta = {"variant": variant, "numpy": np.__version__, "scipy": scipy.__version__,
            "torch": torch.__version__, "NPY_DISABLE_CPU_FEATURES": os.environ.get("NPY_DISABLE_CPU_FEATURES", "")}

    # ---- A. sample_ancestral_index step replay -------------------------------------------
    for i, (name, B, K, gen) in enumerate(step_cases()):
        rng = np.random.default_rng(1000 + i)
        lw = gen(rng, B, K)
        np.random.seed(77 + i)
        u = np.random.uniform(size=[B, 1])  # what inference.py:250 will draw
        np.random.seed(77 + i)
        idx = aesmc.inference.sample_ancestral_index(torch.from_numpy(lw)).numpy()
        with np.errstate(all="ignore"):
            lse = scipy.special.logsumexp(lw, axis=1, keepdims=True)  # the call at math.py:22
            w = aesmc.math.exponentiate_and_normalize(lw, dim=1)
        out["step/%s/lw" % name] = lw
        out["step/%s/u" % name] = u.reshape(B)
        out["step/%s/idx" % name] = idx.astype(np.int32)
        out["step/%s/lse" % name] = lse.reshape(B).astype(np.float32)
        if B * K <= 4096:
            out["step/%s/w" % name] = w.astype(np.float32)

    # ---- B. full infer() traces with the reference's own LGSSM test model ------------------
    def run_infer(tag, algo, B, K, T, seed):
        torch.manual_seed(seed)
        init = lgssm.Initial(0.0, 1.0)
        trans = lgssm.Transition(0.9, 1.0)
        emis = lgssm.Emission(1.0, 0.5)
        prop = lgssm.Proposal(0.8, 0.7)
        lat, obs = aesmc.statistics.sample_from_prior(init, trans, emis, T, B)
        obs = [o.detach() for o in obs]
        np.random.seed(seed)
        u = np.random.uniform(size=[max(T - 1, 1), B])  # same stream as T-1 draws of [B,1]
        np.random.seed(seed)
        torch.manual_seed(seed + 1)
        smc = algo == "smc"
        with torch.no_grad():
            res = aesmc.inference.infer(algo, obs, init, trans, emis, prop, K,
                                        return_log_marginal_likelihood=True, return_latents=True,
                                        return_original_latents=smc, return_log_weight=True,
                                        return_log_weights=True, return_ancestral_indices=smc)
        p = "infer/%s/" % tag
        out[p + "obs"] = torch.stack(obs).numpy()
        out[p + "u"] = u[:T - 1]
        out[p + "log_weights"] = torch.stack(res["log_weights"]).numpy()
        out[p + "log_weight"] = res["log_weight"].numpy()
        out[p + "lml"] = res["log_marginal_likelihood"].numpy()
        out[p + "latents"] = torch.stack(res["latents"]).numpy()
        if smc:
            out[p + "original_latents"] = torch.stack(res["original_latents"]).numpy()
            anc = res["ancestral_indices"]
            out[p + "ancestral_indices"] = (torch.stack(anc).numpy().astype(np.int32) if anc
                                            else np.zeros((0, B, K), np.int32))
        out[p + "params"] = np.array([0.0, 1.0, 0.9, 1.0, 1.0, 0.5, 0.8, 0.7])
        out[p + "proposal_state"] = np.concatenate([v.detach().numpy().ravel() for v in prop.parameters()])
        out[p + "seed"] = np.array([seed])
        # loss value through the reference's losses.get_loss on the same seeds
        np.random.seed(seed)
        torch.manual_seed(seed + 1)
        loss = aesmc.losses.get_loss(obs, K, "aesmc" if smc else "iwae", init, trans, emis, prop)
        out[p + "loss"] = np.array([loss.item()], np.float32)

    run_infer("c1_smc", "smc", 1, 100, 50, 11)       # BASELINE config 1
    run_infer("small_smc", "smc", 3, 64, 12, 12)
    run_infer("small_is", "is", 3, 64, 12, 13)
    run_infer("t1_smc", "smc", 2, 8, 1, 14)

    # ---- C. statistics / math known values computed by the reference ----------------------
    rng = np.random.default_rng(5)
    lw = torch.from_numpy((rng.standard_normal((7, 33)) * 3).astype(np.float32))
    val = torch.from_numpy(rng.standard_normal((7, 33, 4)).astype(np.float32))
    out["stats/lw"] = lw.numpy()
    out["stats/value"] = val.numpy()
    out["stats/log_ess"] = aesmc.statistics.log_ess(lw).numpy()
    out["stats/ess"] = aesmc.statistics.ess(lw).numpy()
    out["stats/mean"] = aesmc.statistics.empirical_mean(val, lw).numpy()
    out["stats/var"] = aesmc.statistics.empirical_variance(val, lw).numpy()
    out["stats/lognormexp_t"] = aesmc.math.lognormexp(lw, dim=1).numpy()
    out["stats/lognormexp_np"] = aesmc.math.lognormexp(lw.numpy(), dim=1)
    out["stats/expnorm_t"] = aesmc.math.exponentiate_and_normalize(lw, dim=1).numpy()
    return out, meta


def main():
    import numpy as np
    if len(sys.argv) == 3 and sys.argv[1] == "--child":
        variant = sys.argv[2]
        out, meta = record(variant)
        np.savez_compressed(os.path.join(HERE, "_tmp_%s.npz" % variant), **out)
        with open(os.path.join(HERE, "_tmp_%s.json" % variant), "w") as f:
            json.dump(meta, f)
        return
    metas = {}
    for variant, disable in (("default", ""), ("avx2", AVX512)):
        env = dict(os.environ)
        if disable:
            env["NPY_DISABLE_CPU_FEATURES"] = disable
        else:
            env.pop("NPY_DISABLE_CPU_FEATURES", None)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", variant], env=env)
        tmp = os.path.join(HERE, "_tmp_%s.npz" % variant)
        data = dict(np.load(tmp))
        os.remove(tmp)
        with open(os.path.join(HERE, "_tmp_%s.json" % variant)) as f:
            metas[variant] = json.load(f)
        os.remove(os.path.join(HERE, "_tmp_%s.json" % variant))
        np.savez_compressed(os.path.join(HERE, "reference_%s.npz" % variant), **data)
        print(variant, "->", len(data), "arrays")
    with open(os.path.join(HERE, "reference_meta.json"), "w") as f:
        json.dump(metas, f, indent=1, sort_keys=True)
    a = np.load(os.path.join(HERE, "reference_default.npz"))
    b = np.load(os.path.join(HERE, "reference_avx2.npz"))
    for k in a.files:
        if k.startswith("step/") and k.endswith("/idx"):
            print(k, "idx differ default vs avx2:", int((a[k] != b[k]).sum()), "of", a[k].size,
                  "| lse rows differ:", int((a[k[:-3] + "lse"].view(np.int32) != b[k[:-3] + "lse"].view(np.int32)).sum()))


if __name__ == "__main__":
    main()
