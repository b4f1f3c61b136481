"""Generate tests/golden/*.npz by running the LIVE reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU
box):  python oracle/gen_golden.py
The fixtures pin the C oracle (oracle/tg_oracle.c) and, through it, the CUDA
path.  Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
OUT = REPO / "tests" / "golden"
REF = "/root/reference"


def main() -> None:
    import torch

    OUT.mkdir(parents=True, exist_ok=True)
    sys.path.insert(0, REF)
    scratch = tempfile.mkdtemp(prefix="tg_golden_")
    os.chdir(scratch)  # the reference mkdirs data_unversioned/ under the cwd (datasets.py:47,177)
    import act  # noqa: E402
    import datasets  # noqa: E402
    import training  # noqa: E402
    import utils  # noqa: E402

    torch.set_num_threads(1)

    # ---------------------------------------------------------------- RNG stream + Categorical
    seeds = [0, 1, 2, 3, 12345]
    rand = []
    for s in seeds:
        torch.manual_seed(s)
        rand.append(torch.rand(700, dtype=torch.float64).numpy())  # crosses one MT19937 twist
    prob_sets = [
        (0.15, 0.7, 0.15),
        (0.1, 0.8, 0.1),
        (0.05, 0.10, 0.70, 0.10, 0.05),
        (0.2, 0.2, 0.2, 0.2, 0.2),
        (1.0, 2.0, 3.0, 2.0, 1.0),  # unnormalised on purpose: Categorical normalises
        (0.3, 0.3, 0.4),
    ]
    cat = {}
    for ip, probs in enumerate(prob_sets):
        n = len(probs)
        values = torch.arange(n) - n // 2
        for s in seeds:
            for length in (4, 9, 16, 64):
                torch.manual_seed(s)
                cat[f"p{ip}_s{s}_n{length}"] = utils.factor_sample(values, torch.tensor(probs), length).numpy().astype(np.int8)
    np.savez_compressed(OUT / "rng.npz", seeds=np.array(seeds), rand=np.stack(rand),
                        prob_sets=np.array([np.pad(p, (0, 5 - len(p))) for p in prob_sets]),
                        prob_lens=np.array([len(p) for p in prob_sets]), **cat)

    # ---------------------------------------------------------------- synthetic demos (utils.py:203-233)
    demo_cfgs = {
        "S4": dict(values=(-1, 0, 1), probs=(0.15, 0.7, 0.15), R=7, S=4, shift=1, n=24),
        "S9": dict(values=(-2, -1, 0, 1, 2), probs=(0.05, 0.10, 0.70, 0.10, 0.05), R=23, S=9, shift=2, n=6),
        "S16": dict(values=(-2, -1, 0, 1, 2), probs=(0.05, 0.10, 0.70, 0.10, 0.05), R=12, S=16, shift=2, n=3),
        "S4u": dict(values=(-2, -1, 0, 1, 2), probs=(0.2, 0.2, 0.2, 0.2, 0.2), R=5, S=4, shift=2, n=8),
    }
    demos = {}
    for name, c in demo_cfgs.items():
        for s in (0, 1, 7):
            torch.manual_seed(s)
            toks, tgts = [], []
            for _ in range(c["n"]):
                a, t = utils.create_synthetic_demo(torch.tensor(c["values"]), torch.tensor(c["probs"]), c["R"], c["S"], c["shift"])
                toks.append(torch.stack(a).numpy())
                tgts.append(t.numpy())
            tail = torch.rand(2, dtype=torch.float64).numpy()  # where the stream stands afterwards
            demos[f"{name}_seed{s}_tokens"] = np.stack(toks).astype(np.int8)
            demos[f"{name}_seed{s}_targets"] = np.stack(tgts).astype(np.int16)
            demos[f"{name}_seed{s}_tail"] = tail
        demos[f"{name}_cfg"] = np.array([c["R"], c["S"], c["shift"], c["n"], len(c["values"])])
        demos[f"{name}_values"] = np.array(c["values"])
        demos[f"{name}_probs"] = np.array(c["probs"], dtype=np.float64)
    # the dataset class runs the same loop through its own method (datasets.py:124-142)
    torch.manual_seed(5)
    ds = datasets.SyntheticDemoDataset(7, 5, 2, 4, "cpu", save_dir=Path(scratch) / "ds_a")
    demos["dataset_seed5_tokens"] = np.stack([torch.stack(torch.load(ds.save_dir / f"action_seq_{i}.pt")).numpy() for i in range(5)]).astype(np.int8)
    demos["dataset_seed5_targets"] = np.stack([torch.load(ds.save_dir / f"target_tensor_{i}.pt").numpy() for i in range(5)]).astype(np.int16)
    np.savez_compressed(OUT / "demos.npz", **demos)

    # ---------------------------------------------------------------- __getitem__ (datasets.py:77-122)
    gi = {}
    for tag, kw, R, dim_t, S in [
        ("a", dict(), 7, 2, 4),
        ("b", dict(), 5, 4, 4),
        ("c", dict(values=(-2, -1, 0, 1, 2), probs=(0.05, 0.10, 0.70, 0.10, 0.05), shift=2), 6, 3, 9),  # Q1: replay uses shift 1
        ("d", dict(), 4, 1, 4),
    ]:
        torch.manual_seed(11)
        ds = datasets.SyntheticDemoDataset(R, 3, dim_t, S, "cpu", save_dir=Path(scratch) / f"gi_{tag}", **kw)
        states, scalars, actions, rewards = [], [], [], []
        for idx in range(len(ds)):
            st, sc, ac, rw = ds[idx]
            states.append(st.numpy()); scalars.append(sc.numpy()); actions.append(ac.numpy()); rewards.append(rw.numpy())
        gi[f"{tag}_cfg"] = np.array([R, dim_t, S, kw.get("shift", 1), 3])
        gi[f"{tag}_tokens"] = np.stack([torch.stack(torch.load(ds.save_dir / f"action_seq_{i}.pt")).numpy() for i in range(3)]).astype(np.int8)
        gi[f"{tag}_targets"] = np.stack([torch.load(ds.save_dir / f"target_tensor_{i}.pt").numpy() for i in range(3)]).astype(np.int16)
        gi[f"{tag}_states"] = np.stack(states).astype(np.int16)
        assert all((np.stack(states) == np.stack(states).astype(np.int16)).ravel())
        gi[f"{tag}_scalars"] = np.stack(scalars).astype(np.float32)
        gi[f"{tag}_actions"] = np.stack(actions).astype(np.int8)
        gi[f"{tag}_rewards"] = np.stack(rewards).astype(np.float32)
        gi[f"{tag}_len"] = np.array(len(ds))
    np.savez_compressed(OUT / "getitem.npz", **gi)

    # ---------------------------------------------------------------- transition (act.py:266-275, utils.py:181-194, training.py:249-268)
    st = {}
    g = torch.Generator().manual_seed(99)
    for S, shift, T in [(4, 1, 2), (9, 2, 3), (16, 2, 2)]:
        nlog = 2 * shift + 1
        k = 8
        n_states = 6
        states = torch.randint(-3, 4, (n_states, T, S, S, S), generator=g).float()
        states[:, 0] *= (torch.rand(n_states, S, S, S, generator=g) < 0.3)  # sparse heads
        actions = torch.randint(0, nlog, (n_states, k, 3 * S), generator=g)
        sparse = torch.rand(n_states, k, 3 * S, generator=g) < 0.6
        actions[sparse] = 1  # NB: token 1 == coefficient 0 under action_to_tensor's fixed shift 1
        actions[:, 0, :S] = 1  # null action: u all zero (under shift 1)
        # make child 1 of state 0 terminal: head := rank-1 tensor of that action
        states[0, 0] = utils.action_to_tensor(actions[0, 1]).float()
        child_heads, null_keep, terminal = [], [], []
        for b in range(n_states):
            s1 = states[b : b + 1]
            children = act.get_child_states(s1, actions[b : b + 1])
            keep = utils.remove_null_actions(s1, children)
            null_keep.append(np.array([i in keep for i in range(k)]))
            terminal.append(np.array([bool(utils.tensor_factorized(utils.get_head_state(c))) for c in children]))
            child_heads.append(torch.cat([c[:, 0] for c in children]).numpy())
            # history shift: slot t of the child is slot t-1 of the parent
            for c in children:
                assert torch.equal(c[:, 1:], s1[:, :-1])
        st[f"S{S}_states"] = states.numpy().astype(np.int16)
        st[f"S{S}_actions"] = actions.numpy().astype(np.int8)
        st[f"S{S}_child_heads"] = np.stack(child_heads).astype(np.int16)
        st[f"S{S}_not_null"] = np.stack(null_keep)
        st[f"S{S}_terminal"] = np.stack(terminal)
        # _take_action arithmetic (tokens - 2, nnz per sample, min over n_samples)
        n_samples = 4
        Bk = 8
        aa = torch.randint(0, 5, (Bk, 1, 3 * S), generator=g)
        aa[torch.rand(Bk, 1, 3 * S, generator=g) < 0.6] = 2
        sb = torch.randint(-2, 3, (Bk, T, S, S, S), generator=g).float()
        sb[:, 0] *= (torch.rand(Bk, S, S, S, generator=g) < 0.2)
        sb[3, 0] = utils.uvw_to_tensor(torch.split(aa[3, 0] - 2, S, dim=-1)).float()  # becomes terminal
        fake = types.SimpleNamespace(
            model=types.SimpleNamespace(fwd_infer=lambda s, c, n_samples=1, _aa=aa: (_aa, None, None)),
            args=types.SimpleNamespace(dim_3d=S, n_samples=n_samples),
        )
        new_sb, new_sc, best, uvw = training.SyntheticDemoTrainingApp._take_action(fake, sb, torch.zeros(Bk, 1))
        st[f"S{S}_ta_states"] = sb.numpy().astype(np.int16)
        st[f"S{S}_ta_tokens"] = aa.numpy().astype(np.int8)
        st[f"S{S}_ta_new_states"] = new_sb.numpy().astype(np.int16)
        st[f"S{S}_ta_scalars"] = new_sc.numpy()
        st[f"S{S}_ta_best_values"] = best.values.numpy()
        st[f"S{S}_ta_best_indices"] = best.indices.numpy()
    np.savez_compressed(OUT / "steps.npz", **st)

    # ---------------------------------------------------------------- Strassen (datasets.py:362-465)
    sd = datasets.StrassenDemoDataset()
    tensor, action_list = datasets.get_strassen_tensor("cpu")
    uu, vv, ww = datasets.get_strassen_factors("cpu")
    np.savez_compressed(
        OUT / "strassen.npz",
        tensor=tensor.numpy().astype(np.int8), action_list=action_list.numpy().astype(np.int8),
        uu=uu.numpy().astype(np.int8), vv=vv.numpy().astype(np.int8), ww=ww.numpy().astype(np.int8),
        states=torch.stack(sd.state_tensor).numpy().astype(np.int8),
        actions=torch.stack(sd.target_action).numpy().astype(np.int8),
        rewards=torch.stack(sd.reward).numpy(), scalars=torch.stack(sd.scalar).numpy(),
        bits=np.array([int(b, 2) for b in sd.bit_info]), n=np.array(len(sd)),
    )

    # ---------------------------------------------------------------- matmul tensor (utils.py:143-161)
    mm = {}
    for n in (2, 3, 4):
        t = utils.build_matmul_tensor(2, n, n, n)
        mm[f"n{n}"] = t.numpy().astype(np.int8)
    np.savez_compressed(OUT / "matmul.npz", **mm)

    # ---------------------------------------------------------------- get_rank (utils.py:134-140), scalars, string keys
    rk = {}
    g = torch.Generator().manual_seed(2024)
    for S in (4, 9, 16):
        n = 40
        T = torch.randint(-2, 3, (n, 1, S, S, S), generator=g).float()
        T *= (torch.rand(n, 1, S, S, S, generator=g) < torch.linspace(0.02, 0.9, n).view(n, 1, 1, 1, 1))
        # low-rank structured cases: sums of few rank-1 terms
        for b in range(0, n, 4):
            acc = torch.zeros(S, S, S)
            for _ in range(1 + b % 5):
                u, v, w = (torch.randint(-2, 3, (S,), generator=g) for _ in range(3))
                acc += utils.uvw_to_tensor((u, v, w))
            T[b, 0] = acc
        rk[f"S{S}_T"] = T[:, 0].numpy().astype(np.int16)
        rk[f"S{S}_rank"] = np.array([utils.get_rank(T[b : b + 1]) for b in range(n)])
    rk["mm_rank"] = np.array([utils.get_rank(utils.build_matmul_tensor(1, n, n, n).unsqueeze(0)) for n in (2, 3, 4)])
    np.savez_compressed(OUT / "ranks.npz", **rk)

    misc = {
        "scalars_b": utils.get_scalars(torch.zeros(5, 2, 4, 4, 4), 3).numpy(),
        "scalars_s": utils.get_scalars(torch.zeros(2, 4, 4, 4), 3, batch_size=False).numpy(),
        "str_key": np.array(utils.state_to_str(torch.tensor([[1.0, -2.0], [0.0, 3.0]]))),
    }
    np.savez_compressed(OUT / "misc.npz", **misc)
    # ---------------------------------------------------------------- MCTS trajectory (act.py:8-64) with a fake model
    sys.path.insert(0, str(REPO))
    from tests.fake_model import FakeAlphaTensor

    mc = {}
    for tag, S, T, max_actions, n_sim, seed in [("a", 4, 2, 4, 6, 0), ("b", 4, 1, 3, 5, 1), ("c", 9, 2, 3, 4, 2)]:
        torch.manual_seed(100 + seed)
        _, start = utils.create_synthetic_demo(torch.tensor((-1, 0, 1)), torch.tensor((0.1, 0.8, 0.1)), 2, S, 1)
        init = torch.cat((start.unsqueeze(0), torch.zeros(T - 1, S, S, S)))
        fm = FakeAlphaTensor(dim_3d=S, n_samples=4, n_logits=3, seed=seed)
        state_seq, policy_seq, reward_seq = act.actor_prediction(fm, init, max_actions, n_sim, 100)
        mc[f"{tag}_cfg"] = np.array([S, T, max_actions, n_sim, seed])
        mc[f"{tag}_init"] = init.numpy().astype(np.int16)
        mc[f"{tag}_states"] = torch.stack(state_seq).numpy().astype(np.int16)
        mc[f"{tag}_policy"] = policy_seq.numpy()
        mc[f"{tag}_rewards"] = reward_seq.numpy()
        mc[f"{tag}_calls"] = np.array(fm.calls)
    np.savez_compressed(OUT / "mcts.npz", **mc)

    # ---------------------------------------------------------------- star-import surface of the three modules
    import json

    surface = {m.__name__: sorted(n for n in dir(m) if not n.startswith("_")) for m in (utils, datasets, act)}
    (OUT / "surface.json").write_text(json.dumps(surface, indent=1))
    print("golden fixtures written to", OUT)
    for f in sorted(OUT.glob("*.npz")):
        print(f"  {f.name}: {f.stat().st_size} B")


if __name__ == "__main__":
    main()
