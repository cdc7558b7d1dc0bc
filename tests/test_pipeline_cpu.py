"""CPU: host logic of the pipeline driver (settings validation, checkpoint file, signal handler)."""
import json
import os
import signal

import pytest

from fava_b200.__main__ import InterruptHandler, Pipeline


def test_settings_validation_and_checkpoint(tmp_path):
    (tmp_path / "x_hdf5_plt_cnt_0000").write_bytes(b"x")
    good = {"data folder": str(tmp_path), "output folder": str(tmp_path), "basename": "x", "dimension": 3, "model": "m"}
    (tmp_path / "pipeline_settings.json").write_text(json.dumps(good))
    p = Pipeline(tmp_path)
    p.restart()
    assert p.model.nfiles(file_type="plt") == 1 and p.half_width == 16e5 and p.length == 32e5
    p.checkpoint_data["reynolds stress"] = {"index": 1}
    p.checkpoint()
    q = Pipeline(tmp_path)
    q.restart()
    assert q.checkpoint_data["reynolds stress"] == {"index": 1} and q.checkpoint_data["settings"] == good
    bad = dict(good)
    bad["dimension"] = "3"
    (tmp_path / "pipeline_settings.json").write_text(json.dumps(bad))
    with pytest.raises(AssertionError):
        Pipeline(tmp_path).restart()


def test_interrupt_handler_calls_checkpoint_once():
    calls = []
    with InterruptHandler(external_handler=lambda: calls.append(1)) as h:
        os.kill(os.getpid(), signal.SIGTERM)
        assert h.interrupted and h.signal == signal.SIGTERM
    assert calls == [1]  # released by the signal, not again on exit
    calls.clear()
    with InterruptHandler(external_handler=lambda: calls.append(1)):
        pass
    assert calls == [1]  # normal exit also checkpoints (reference _mpi.py:118-119)
    assert signal.getsignal(signal.SIGTERM) is not None
