"""Model clock (reference: src/fluids2d/timeline.py:4-51).  Host bookkeeping."""


def _two_sum_step(total, carry, increment):
    """one step of compensated summation: returns (new total, new carry)"""
    corrected = increment - carry
    new_total = total + corrected
    return new_total, (new_total - total) - corrected


class Time:
    """`t`, `ite`, `dt` and the cadence predicates the run loop polls
    (model.py:45-51).  `t0` / `ite0` are the origin of the current run()."""

    def __init__(self, param):
        self.param = param
        self.t, self.ite = 0.0, 0
        self.t0, self.ite0 = 0, 0
        self.dt = 0.01 if param.dt <= 0 else param.dt
        self._c = 0.0     # rounding error carried by the compensated sum

    def pushforward(self):
        # t advances by Kahan summation so that ten steps of 0.1 land on 1.0
        # exactly, as in the reference (timeline.py:20-35)
        self.t, self._c = _two_sum_step(self.t, self._c, self.dt)
        self.ite += 1

    def _every(self, n):
        return self.ite % n == 0 or self.finished

    @property
    def finished(self):
        out_of_time = self.t >= self.t0 + self.param.tend
        out_of_steps = self.ite >= self.ite0 + self.param.maxite
        return out_of_time or out_of_steps

    @property
    def update_anim(self):
        return bool(self.param.animation) and self._every(self.param.nplot)

    @property
    def save_to_file(self):
        return self.param.nhis != 0 and self._every(self.param.nhis)

    def tostring(self):
        return f"t={self.t:.2f}"
