"""Oracle-backed stand-in for an EINCM_FLAG_EVENT_SPLIT plan (CPU tensors), so that the collective sequence of
eincm_b200.parallel.EventSplitObjective can be exercised over gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

from oracle import eincm_oracle as O


FIX_SCALE = float(2 ** 21) * 2.0 * np.pi          # the library's fixed-point images: 2^21 * 2 pi * value (include/eincm.h)


class OracleSplitPlan:
    def __init__(self, sensor_size):
        self.sensor_size = tuple(sensor_size)
        self.rank, self.world = 0, 1
        self.fixed = False

    def set_split_fixed_point(self, on=True):
        """The collective then runs on int64 images (iwe_fix); backward() takes the complete image from there."""
        self.fixed = bool(on)

    def set_event_split(self, rank, world):
        self.rank, self.world = rank, world

    def set_window(self, xs, ys, ts, edges, edge_ts):
        self.xs, self.ys, self.ts = np.asarray(xs), np.asarray(ys), np.asarray(ts, dtype=np.float64)
        self.edges, self.edge_ts = np.asarray(edges), np.asarray(edge_ts, dtype=np.float64)
        self._zero = torch.from_numpy(O.events_to_pdf_frame(self.xs, self.ys, self.sensor_size))
        self._mask = torch.from_numpy(O.make_event_mask(self.xs, self.ys, self.sensor_size).astype(np.uint8))
        self.final = False

    def zero_iwe(self):
        return self._zero

    def event_mask(self):
        return self._mask

    def window_finalize(self):
        self.final = True

    def forward_events(self, theta, hp):
        assert self.final
        self.theta = theta.numpy()
        self._iwe = torch.from_numpy(O.partial_images(self.theta, self.xs, self.ys, self.ts, self.edge_ts, self.sensor_size))
        if self.fixed:            # quantised like the library's votes (here per cell, not per vote: a stand-in for the SEQUENCE, not the kernels)
            self._fix = torch.from_numpy(np.rint(self._iwe.numpy() * FIX_SCALE).astype(np.int64))

    def iwe(self):
        assert not self.fixed, 'fixed-point split: the caller must all-reduce iwe_fix(), not iwe()'
        return self._iwe

    def iwe_fix(self):
        assert self.fixed
        return self._fix

    def backward(self, hp, loss_out, grad_out):
        iwe = self._fix.numpy().astype(np.float64) / FIX_SCALE if self.fixed else self._iwe.numpy()
        loss, grad = O.split_value_and_grad(self.theta, iwe, self._zero.numpy(), self._mask.numpy().astype(bool),
                                            self.xs, self.ys, self.ts, self.edges, self.edge_ts, hp['alpha'], hp['beta'], hp['gamma'],
                                            hp['delta'], hp['cur_pyr_lvl'], 5, self.sensor_size,
                                            include_replicated_grad=(self.rank == 0))
        loss_out[0] = loss
        grad_out.copy_(torch.from_numpy(grad))
