"""Driver plugins.

The plugin surface of the reference is kept as is: a driver is any object with
`process_lidar(ranges) -> (speed, steering_angle)` (or `process_lidar(ranges, state)`), see
drivers/template.py:1-18 and ft_grandprix/custom.py:1398-1411.  `Fleet.drive_host` runs such
objects unchanged.  This module adds the batched variants of the two bundled drivers

    BatchedNidcDriver  <- ft_grandprix/nidc.py:3-131
    BatchedFastDriver  <- ft_grandprix/fast.py:3-139

as plain torch tensor programs: `process_lidar(ranges[n, 90]) -> (speed[n], steering[n])` on
whatever device `ranges` lives on.  They reproduce the sequential semantics of the Python loops
(SURVEY Appendix D): disparities are found once on the unmodified scan, then extended one
after the other, each reading what earlier ones wrote; np.argmax first-index ties.  The CUDA
kernel behind `Fleet.drive()` (csrc/fleet.cu) is the fast path; these are the reference-shaped
variant users can subclass.
"""
import math

import torch


class LobotomyDriver:
    """ft_grandprix/lobotomy.py:1-3"""
    def process_lidar(self, ranges):
        return 0.0, 0.0


class BatchedLobotomyDriver:
    def process_lidar(self, ranges):
        z = torch.zeros(ranges.shape[0], dtype=torch.float64, device=ranges.device)
        return z, z.clone()


class BatchedNidcDriver:
    CAR_WIDTH = 0.12                 # nidc.py:5
    DIFFERENCE_THRESHOLD = 0.6       # nidc.py:7
    SPEED = 0.5                      # nidc.py:8
    SAFETY_PERCENTAGE = 300.         # nidc.py:10

    def _extend(self, ranges):
        r = ranges.to(torch.float64)
        n, nb = r.shape
        rpp = (2 * math.pi) / nb                                     # nidc.py:121
        eighth = int(nb / 8)                                         # nidc.py:17
        proc = r[:, eighth:nb - eighth].clone()                      # nidc.py:18
        m = proc.shape[1]
        diff = torch.zeros_like(proc)
        diff[:, 1:] = (proc[:, 1:] - proc[:, :-1]).abs()             # nidc.py:26-29
        disp = diff > self.DIFFERENCE_THRESHOLD                      # nidc.py:31-40 (computed once)
        width = (self.CAR_WIDTH / 2) * (1 + self.SAFETY_PERCENTAGE / 100)   # nidc.py:93
        idx = torch.arange(m, device=r.device).unsqueeze(0)          # [1, m]
        raised = torch.zeros(n, dtype=torch.bool, device=r.device)
        for i in range(1, m):                                        # nidc.py:94-104, in index order
            act = disp[:, i] & ~raised
            a, b = proc[:, i - 1], proc[:, i]
            # np.argmin / np.argmax on 2 elements: first NaN wins, ties -> first
            amin = torch.where(a.isnan(), 0, torch.where(b.isnan(), 1, (b < a).long()))
            amax = torch.where(a.isnan(), 0, torch.where(b.isnan(), 1, (b > a).long()))
            close = i - 1 + amin
            dist = torch.where(amin.bool(), b, a)
            num = torch.ceil(2 * torch.atan(width / (2 * dist)) / rpp)      # nidc.py:57-58
            bad = num.isnan() & act
            raised |= bad                                            # int(nan) raises -> ctrl kept
            act = act & ~bad
            num = torch.nan_to_num(num, nan=0.0)
            right = (amin < amax).unsqueeze(1)
            c = close.unsqueeze(1); k = num.unsqueeze(1)
            cover = torch.where(right, (idx > c) & (idx <= c + k), (idx < c) & (idx >= c - k))
            cover &= act.unsqueeze(1) & (proc > dist.unsqueeze(1))
            proc = torch.where(cover, dist.unsqueeze(1), proc)
        # np.argmax: first maximum; a NaN anywhere wins at its first position
        nan_any = proc.isnan().any(1)
        first_nan = proc.isnan().float().argmax(1)
        mx = torch.nan_to_num(proc, nan=-math.inf).max(1, keepdim=True).values
        first_max = (torch.nan_to_num(proc, nan=-math.inf) == mx).float().argmax(1)
        best = torch.where(nan_any, first_nan, first_max)
        ang = (best.to(torch.float64) - m / 2.0) * rpp               # nidc.py:112
        steer = ang.clamp(-math.radians(90), math.radians(90))       # nidc.py:113
        return steer, raised, r

    def process_lidar(self, ranges):
        steer, raised, _ = self._extend(ranges)
        speed = self.SPEED * 5 * (1 - steer.abs() / (1.57 * 2))      # nidc.py:130
        self.raised = raised
        return speed, steer


class BatchedFastDriver(BatchedNidcDriver):
    CAR_WIDTH = 0.06                 # fast.py:4

    def process_lidar(self, ranges):
        steer, raised, r = self._extend(ranges)
        slow = torch.clamp(self.SPEED * 5 * (1 - steer.abs() / math.pi), max=2.0)     # fast.py:138
        speed = torch.where((steer.abs() < 0.1) & (r[:, 0] > 0.5), torch.full_like(slow, 7.0), slow)  # fast.py:135-136
        self.raised = raised
        return speed, steer
