"""Oracle-backed stand-in for an EINCM_FLAG_EVENT_SPLIT plan (CPU tensors), so that the collective sequence of
eincm_b200.parallel.EventSplitObjective can be exercised over gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

from oracle import eincm_oracle as O


class OracleSplitPlan:
    def __init__(self, sensor_size):
        self.sensor_size = tuple(sensor_size)
        self.rank, self.world = 0, 1

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

    def iwe(self):
        return self._iwe

    def backward(self, hp, loss_out, grad_out):
        loss, grad = O.split_value_and_grad(self.theta, self._iwe.numpy(), self._zero.numpy(), self._mask.numpy().astype(bool),
                                            self.xs, self.ys, self.ts, self.edges, self.edge_ts, hp['alpha'], hp['beta'], hp['gamma'],
                                            hp['delta'], hp['cur_pyr_lvl'], 5, self.sensor_size,
                                            include_replicated_grad=(self.rank == 0))
        loss_out[0] = loss
        grad_out.copy_(torch.from_numpy(grad))
