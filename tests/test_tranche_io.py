"""CPU tests of the tranche file formats (reference CSV + binary row-major) and their ordering rules."""
import datetime as dt
import os

import numpy as np
import pandas as pd
import pytest

import bodywork_mlops_demo_b200 as b2
from bodywork_mlops_demo_b200 import stage_1_train_model as s1
from bodywork_mlops_demo_b200 import tranche_io as tio
from oracle import ols_oracle as orc


def test_binary_tranche_round_trip(tmp_path):
    X, y = orc.generate_dataset(1001, 7, seed=1, dtype=np.float32)
    path = tmp_path / "regression-dataset-2021-04-08.b2t"
    tio.write_tranche(str(path), X, y, dt.date(2021, 4, 8))
    assert os.path.getsize(path) == 64 + (1001 * 7 * 4 + 63) // 64 * 64 + 1001 * 4
    assert tio.read_header(str(path)) == (1001, 7, "f32", dt.date(2021, 4, 8))
    X2, y2, kind, day = tio.read_tranche(str(path))
    assert np.array_equal(X, X2) and np.array_equal(y, y2) and kind == "f32"
    # read straight into caller buffers (what the pinned-memory path does)
    bx, by = np.empty((1001, 7), np.float32), np.empty(1001, np.float32)
    tio.read_tranche(str(path), bx, by)
    assert np.array_equal(bx, X)
    with pytest.raises(ValueError):
        tio.read_tranche(str(path), np.empty((5, 7), np.float32), by)
    bits = b2.native.to_bf16_bits(X)
    p2 = tmp_path / "regression-dataset-2021-04-09.b2t"
    tio.write_tranche(str(p2), bits, y, dt.date(2021, 4, 9), x_kind="bf16")
    assert tio.read_header(str(p2))[2] == "bf16" and np.array_equal(tio.read_tranche(str(p2))[0], bits)
    (tmp_path / "junk.b2t").write_bytes(b"nope")
    with pytest.raises(RuntimeError):
        tio.read_header(str(tmp_path / "junk.b2t"))


def test_mixed_csv_and_binary_tranches_load_in_date_order(tmp_path):
    folder = tmp_path / "datasets"
    folder.mkdir()
    Xa, ya = orc.generate_dataset(300, 3, seed=2, dtype=np.float32)
    Xb, yb = orc.generate_dataset(200, 3, seed=3, dtype=np.float32)
    tio.write_tranche(str(folder / "regression-dataset-2021-04-09.b2t"), Xb, yb, dt.date(2021, 4, 9))
    df = pd.DataFrame(Xa, columns=["X0", "X1", "X2"]); df["y"] = ya; df["date"] = "2021-04-08"
    df.to_csv(folder / "regression-dataset-2021-04-08.csv", index=False)
    items = tio.list_tranches(str(folder))
    assert [d for _, d in items] == [dt.date(2021, 4, 8), dt.date(2021, 4, 9)]
    X, y, newest, keep = tio.load_all(str(folder))
    assert newest == dt.date(2021, 4, 9) and keep is None
    np.testing.assert_allclose(X, np.concatenate([Xa, Xb]), rtol=1e-6)
    np.testing.assert_allclose(y, np.concatenate([ya, yb]), rtol=1e-6)
    # the stage's own loader sees the same rows (DataFrame form, reference signature)
    data, newest2 = s1.download_latest_dataset(str(tmp_path))
    assert newest2 == newest and len(data) == 500 and s1.feature_columns(data) == ["X0", "X1", "X2"]
    np.testing.assert_allclose(data[["X0", "X1", "X2"]].to_numpy(np.float32), X, rtol=1e-6)


def test_state_file_round_trip(tmp_path):
    from bodywork_mlops_demo_b200 import incremental
    S = orc.gram_stats(*orc.generate_dataset(100, 4, seed=5))
    incremental.save_state(S, dt.date(2021, 4, 8), str(tmp_path))
    assert np.array_equal(incremental.load_state(dt.date(2021, 4, 8), str(tmp_path)), S)
