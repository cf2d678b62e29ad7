"""CPU: the data formats either side of the hot path (red-diffeq_b200/utils/io.py): .npy families in, per-model .npz out,
initial models -- checked against the reference's conventions (scripts/run_inversion.py:143-216, utils/data_trans.py:66-102)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _family(tmp_path, n=5, ns=2, nt=9, nrec=6, nz=6, nx=6):
    rng = np.random.default_rng(0)
    seis = rng.standard_normal((n, ns, nt, nrec)).astype(np.float32)
    vel = (1500 + 3000 * rng.random((n, 1, nz, nx))).astype(np.float32)
    np.save(tmp_path / "seis.npy", seis)
    np.save(tmp_path / "vel.npy", vel)
    return seis, vel


def test_family_batches_and_loading(tmp_path):
    from red_diffeq_b200.utils.io import Family
    seis, vel = _family(tmp_path)
    fam = Family(tmp_path / "seis.npy", tmp_path / "vel.npy")
    assert len(fam) == 5 and isinstance(fam.seismic, np.memmap)
    assert fam.batches(2) == [(0, 2), (2, 4), (4, 5)] and fam.batches(25) == [(0, 5)]
    assert fam.batches(2, sample_index=3) == [(3, 4)]
    with pytest.raises(IndexError):
        fam.batches(2, sample_index=5)
    s, v = fam.load_batch(2, 4, "cpu")
    assert s.dtype == torch.float32 and np.array_equal(s.numpy(), seis[2:4]) and np.array_equal(v.numpy(), vel[2:4])
    fam.check_against(dict(ns=2, nt=9, ng=6))
    with pytest.raises(ValueError):
        fam.check_against(dict(ns=2, nt=10, ng=6))
    np.save(tmp_path / "bad.npy", vel[:3])
    with pytest.raises(ValueError):
        Family(tmp_path / "seis.npy", tmp_path / "bad.npy")


def test_initial_models_follow_the_reference():
    from scipy.ndimage import gaussian_filter
    from red_diffeq_b200 import v_normalize
    from red_diffeq_b200.utils.io import initial_batch, prepare_initial_model
    rng = np.random.default_rng(1)
    v = torch.tensor((1500 + 3000 * rng.random((1, 1, 8, 10))).astype(np.float32))
    sm = prepare_initial_model(v, "smoothed", sigma=2.0)
    assert np.array_equal(sm.numpy(), gaussian_filter(v_normalize(v.numpy()), sigma=2.0).astype(np.float32))
    hom = prepare_initial_model(v, "homogeneous")
    assert float(hom.min()) == float(hom.max()) == float(v_normalize(v.numpy())[0, 0, 0].min())
    lin = prepare_initial_model(v, "linear").numpy()
    assert lin.shape == (1, 1, 8, 10) and np.all(np.diff(lin[0, 0, :, 0]) > 0) and np.all(lin[0, 0, :, 0:1] == lin[0, 0])
    with pytest.raises(AssertionError):
        prepare_initial_model(v, "random")
    batch = initial_batch(torch.cat([v, v * 0.9 + 200]), "smoothed", 2.0)
    assert batch.shape == (2, 1, 10, 12) and float(batch[:, :, 0].abs().max()) == 0.0 and torch.equal(batch[:1, :, 1:-1, 1:-1], sm)


def test_results_round_trip(tmp_path):
    from red_diffeq_b200.utils.io import RESULT_KEYS, save_batch_results
    B, ts = 2, 4
    mu = torch.rand(B, 1, 6, 6)
    init = torch.rand(B, 1, 8, 8)
    vel = 1500 + 3000 * torch.rand(B, 1, 6, 6)
    res = [{k: [np.float32(t + 10 * i) for t in range(ts)] for k in RESULT_KEYS[3:]} for i in range(B)]
    paths = save_batch_results(7, 9, mu, res, init, vel, tmp_path / "out" / "family")
    assert [p.split("/")[-1] for p in paths] == ["7_results.npz", "8_results.npz"]
    z = np.load(paths[1])
    assert sorted(z.files) == sorted(RESULT_KEYS)
    assert np.array_equal(z["result"], mu[1, 0].numpy()) and np.array_equal(z["ground_truth"], vel[1, 0].numpy())
    assert np.array_equal(z["initial_velocity"], init[1, 0, 1:-1, 1:-1].numpy())
    assert z["mae"].shape == (ts,) and z["mae"][2] == 12.0
