"""SURVEY 8(f) rows 3 and 4 against fixtures drawn from the live reference (oracle/gen_golden_r2.py): the batched
played-game return rule (act.py:59-62) and the replay buffers as device ring buffers with the reference's mixture
sampling (datasets.py:161-359)."""
import numpy as np
import pytest
import torch

from tests.helpers import dense_to_slab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S", [4, 9])
def test_batched_episode_returns_match_reference(golden, S):
    from mat_mul_b200 import env

    g = golden["returns"]
    heads, lens, want_seq, want_end = g[f"S{S}_heads"], g[f"S{S}_lens"], g[f"S{S}_reward_seq"], g[f"S{S}_end"]
    slab = torch.from_numpy(dense_to_slab(heads.astype(np.int32))).cuda()
    steps = torch.from_numpy(lens.astype(np.int32)).cuda()
    returns, seq, ranks = env.episode_returns(slab, steps, S, max_len=want_seq.shape[1])
    assert np.array_equal(-ranks.cpu().numpy(), want_end)  # -get_rank(final state), exact rank == the reference's SVD rank
    assert np.array_equal(seq.cpu().numpy(), want_seq)
    assert np.array_equal(returns.cpu().numpy(), want_seq[np.arange(len(lens)), lens - 1])


def test_played_games_ring_matches_reference(golden, tmp_path, monkeypatch):
    from mat_mul_b200 import datasets as ds

    monkeypatch.chdir(tmp_path)
    g = golden["buffers"]
    pg = ds.PlayedGamesDataset(3, "cpu")
    for gi in range(5):
        states = [torch.from_numpy(s.astype(np.float32)) for s in g[f"g{gi}_states"]]
        pg.add_game(states, torch.from_numpy(g[f"g{gi}_policy"]), torch.from_numpy(g[f"g{gi}_reward"]))
        assert pg._states.is_cuda and pg._states.dtype == torch.int16  # the ring lives in HBM
        assert len(pg) == int(g[f"after{gi}_len"]) and pg.game_pointer == int(g[f"after{gi}_pointer"])
        items = [pg[i] for i in range(len(pg))]
        assert np.array_equal(np.stack([it[0].numpy() for it in items]), g[f"after{gi}_state"].astype(np.float32))
        assert np.array_equal(np.stack([it[1].numpy() for it in items]), g[f"after{gi}_scalar"])
        assert np.array_equal(np.stack([it[2].numpy() for it in items]), g[f"after{gi}_action"])
        assert np.array_equal(np.stack([it[3].numpy() for it in items]), g[f"after{gi}_reward"])
        assert items[0][0].dtype == torch.float32 and items[0][2].dtype == torch.int64 and items[0][3].dtype == torch.int64
        batch = pg.__getitems__(list(range(len(pg))))
        assert all(torch.equal(a, b) for x, y in zip(batch, items) for a, b in zip(x, y))
    # the caller's lists are snapshotted: later mutation does not change the buffer
    before = pg[0][0].clone()
    states[0].add_(5)
    assert torch.equal(pg[0][0], before)


def test_tensor_game_dataset_mixture_matches_reference(golden, tmp_path, monkeypatch):
    from mat_mul_b200 import datasets as ds

    monkeypatch.chdir(tmp_path)
    g = golden["buffers"]
    tg = ds.TensorGameDataset(50, 0.7, 3, 2, 4, "cpu")
    games = [([torch.from_numpy(s.astype(np.float32)) for s in g[f"g{gi}_states"]], torch.from_numpy(g[f"g{gi}_policy"]),
              torch.from_numpy(g[f"g{gi}_reward"])) for gi in range(5)]
    for game in games[:4]:
        tg.add_played_game(*game)
    tg.add_best_game(*games[4])
    for case, (fs, fb) in enumerate(((0.7, 0.0), (0.5, 0.2), (0.25, 0.05))):
        tg.set_fractions(fs, fb)
        torch.manual_seed(100 + case)
        np.random.seed(200 + case)
        tg.resample_buffer_indexes()
        assert np.array_equal(tg.is_synth.numpy(), g[f"mix{case}_is_synth"])
        assert np.array_equal(tg.index_synth.numpy(), g[f"mix{case}_index_synth"])
        assert np.array_equal(tg.index_played.numpy(), g[f"mix{case}_index_played"])
        if fb > 0:
            assert np.array_equal(tg.index_best.numpy(), g[f"mix{case}_index_best"])
        kinds, inner = g[f"mix{case}_kind"], g[f"mix{case}_inner"]
        items = tg.__getitems__(list(range(len(tg))))  # one launch for the synthetic share, one gather per replay buffer
        other_states, other_rewards = [], []
        for idx in range(len(tg)):
            buf, j = tg._route(idx)
            assert (buf is tg.buffer_synth) == (kinds[idx] == 0)
            if kinds[idx] == 0:
                assert j == inner[idx]
            else:
                other_states.append(items[idx][0].numpy())
                other_rewards.append(items[idx][3].numpy())
            single = tg[idx]
            assert all(torch.equal(a, b) for a, b in zip(single, items[idx]))
        if other_states:
            assert np.array_equal(np.stack(other_states), g[f"mix{case}_items_state"].astype(np.float32))
            assert np.array_equal(np.stack(other_rewards), g[f"mix{case}_items_reward"])
