"""Round-2 fixtures from the LIVE reference (/root/reference): run in the build container only.
  python oracle/gen_golden_r2.py
returns.npz  the played-game return rule of actor_prediction (act.py:59-62) on batches of final states
buffers.npz  PlayedGamesDataset ring semantics and TensorGameDataset's mixture sampling (datasets.py:161-359)
Nothing here is imported by the product."""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
OUT = REPO / "tests" / "golden"
REF = "/root/reference"


def main() -> None:
    import torch

    sys.path.insert(0, REF)
    os.chdir(tempfile.mkdtemp(prefix="tg_golden_r2_"))  # the reference mkdirs data_unversioned/ under the cwd
    import datasets  # noqa: E402
    import utils  # noqa: E402

    torch.set_num_threads(1)
    rng = np.random.default_rng(2)

    # ---------------------------------------------------------------- returns (act.py:59-62)
    ret = {}
    for S in (4, 9):
        B, K = 40, 12
        heads = (rng.integers(-2, 3, (B, S, S, S)) * (rng.random((B, S, S, S)) < 0.15)).astype(np.int16)
        heads[::5] = 0  # solved games
        heads[1] = np.einsum("i,j,k->ijk", *[rng.integers(-1, 2, S) for _ in range(3)])  # a rank-one residual
        lens = rng.integers(1, K + 1, B)
        seqs = np.zeros((B, K), dtype=np.int64)
        ends = np.zeros(B, dtype=np.int64)
        for b in range(B):
            state = torch.zeros(1, 2, S, S, S)
            state[0, 0] = torch.from_numpy(heads[b].astype(np.float32))
            n = int(lens[b])  # len(policy_seq): actions played
            end_state_reward = -utils.get_rank(state)                                           # act.py:59
            reward_seq = torch.cumsum(torch.tensor([-1] * (n - 1) + [-1 + end_state_reward]), dim=0)  # act.py:60-62
            seqs[b, :n] = reward_seq.numpy()
            ends[b] = end_state_reward
        ret[f"S{S}_heads"], ret[f"S{S}_lens"], ret[f"S{S}_reward_seq"], ret[f"S{S}_end"] = heads, lens, seqs, ends
    np.savez_compressed(OUT / "returns.npz", **ret)

    # ---------------------------------------------------------------- replay buffers (datasets.py:161-359)
    S, T, n_steps, n_logits = 4, 2, 12, 3
    buf = {}
    pg = datasets.PlayedGamesDataset(3, "cpu")  # ring of 3 games
    games = []
    for gi, n in enumerate((2, 4, 3, 1, 2)):  # five games into a ring of three: slots 0, 1 are overwritten
        states = [torch.from_numpy((rng.integers(-2, 3, (T, S, S, S)) * (rng.random((T, S, S, S)) < 0.3)).astype(np.float32)) for _ in range(n)]
        pol = torch.from_numpy(rng.random((n, n_steps, n_logits)).astype(np.float32))
        rew = torch.cumsum(torch.tensor([-1] * (n - 1) + [-1 - int(rng.integers(0, 5))]), dim=0)
        games.append((states, pol, rew))
        pg.add_game(states, pol, rew)
        buf[f"g{gi}_states"] = np.stack([s.numpy() for s in states]).astype(np.int8)
        buf[f"g{gi}_policy"], buf[f"g{gi}_reward"] = pol.numpy(), rew.numpy()
        items = [pg[i] for i in range(len(pg))]
        buf[f"after{gi}_len"] = np.array(len(pg))
        buf[f"after{gi}_pointer"] = np.array(pg.game_pointer)
        buf[f"after{gi}_state"] = np.stack([it[0].numpy() for it in items]).astype(np.int8)
        buf[f"after{gi}_scalar"] = np.stack([it[1].numpy() for it in items])
        buf[f"after{gi}_action"] = np.stack([it[2].numpy() for it in items])
        buf[f"after{gi}_reward"] = np.stack([it[3].numpy() for it in items])
    # mixture sampling: index arrays the reference draws (numpy + torch global generators) for given buffer lengths
    tg = datasets.TensorGameDataset(50, 0.7, 3, T, S, "cpu")
    for states, pol, rew in games[:4]:
        tg.add_played_game(states, pol, rew)
    tg.add_best_game(*games[4])
    for case, (fs, fb) in enumerate(((0.7, 0.0), (0.5, 0.2), (0.25, 0.05))):
        tg.set_fractions(fs, fb)
        torch.manual_seed(100 + case)
        np.random.seed(200 + case)
        tg.resample_buffer_indexes()
        buf[f"mix{case}_is_synth"] = tg.is_synth.numpy()
        buf[f"mix{case}_index_synth"] = tg.index_synth.numpy()
        buf[f"mix{case}_index_played"] = tg.index_played.numpy()
        buf[f"mix{case}_index_best"] = tg.index_best.numpy() if tg.index_best is not None and fb > 0 else np.zeros(0, dtype=np.int64)
        # which buffer / inner index serves every dataset index, and the non-synthetic items themselves
        kinds, inner = [], []
        for idx in range(len(tg)):
            if tg.is_synth[idx]:
                kinds.append(0); inner.append(int(tg.index_synth[int(tg.is_synth[:idx].sum())]))
            else:
                st, sc, ac, rw = tg[idx]
                kinds.append(1); inner.append(-1)
                buf.setdefault(f"mix{case}_items_state", []).append(st.numpy().astype(np.int8))
                buf.setdefault(f"mix{case}_items_reward", []).append(rw.numpy())
        buf[f"mix{case}_kind"], buf[f"mix{case}_inner"] = np.array(kinds), np.array(inner)
        for k in (f"mix{case}_items_state", f"mix{case}_items_reward"):
            buf[k] = np.stack(buf[k]) if k in buf else np.zeros(0)
    np.savez_compressed(OUT / "buffers.npz", **buf)
    print("wrote returns.npz, buffers.npz")


if __name__ == "__main__":
    main()
