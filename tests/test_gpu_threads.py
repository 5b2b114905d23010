"""The reference documents apply as thread-safe on a shared filter (include/iterative/savgolFilter.h:16-19).
Several host threads apply the same filter concurrently, each on its own CUDA stream, device and
host pointers mixed."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402


def test_concurrent_apply_on_shared_filter(oracle):
    f = sg.SavgolFilter(12, 4, 0, 1.0, "polynomial")
    f2 = sg.Savgol2DFilter(3, 3, 2)
    o = oracle.Filter1D(12, 4, 0, 1.0, "polynomial")
    o2 = oracle.Filter2D(3, 3, 2)
    rng = np.random.default_rng(5)
    xs = [rng.standard_normal((7, 3000 + 17 * i)).astype(np.float32) for i in range(6)]
    imgs = [rng.standard_normal((90 + i, 130)).astype(np.float32) for i in range(6)]
    errs = []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for rep in range(8):
                    if (i + rep) % 3 == 0:
                        y = f.apply(xs[i])                                   # host pointers (staged)
                    else:
                        y = f.apply(torch.from_numpy(xs[i]).cuda()).cpu().numpy()
                    assert np.max(np.abs(y - o.apply(xs[i]))) <= 1e-6 * float(np.abs(xs[i]).max())
                    z = f2.apply(torch.from_numpy(imgs[i]).cuda(), "reflect").cpu().numpy()
                    assert np.max(np.abs(z - o2.apply(imgs[i], "reflect"))) <= 1e-6 * float(np.abs(imgs[i]).max())
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
