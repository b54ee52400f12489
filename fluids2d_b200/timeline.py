"""Model clock (reference: src/fluids2d/timeline.py:4-51).  Host bookkeeping."""


class Time:
    def __init__(self, param):
        self.param = param
        self.t = 0.0
        self.ite = 0
        self.t0 = 0
        self.ite0 = 0
        self.dt = param.dt if param.dt > 0 else 0.01
        self._c = 0.0     # Kahan compensation

    @property
    def finished(self):
        return (self.t >= self.t0 + self.param.tend) or (self.ite >= self.ite0 + self.param.maxite)

    def pushforward(self):
        # compensated t += dt (timeline.py:20-35)
        y = self.dt - self._c
        t = self.t + y
        self._c = (t - self.t) - y
        self.t = t
        self.ite += 1

    def tostring(self):
        return f"t={self.t:.2f}"

    @property
    def update_anim(self):
        return self.param.animation and ((self.ite % self.param.nplot == 0) or self.finished)

    @property
    def save_to_file(self):
        if self.param.nhis == 0:
            return False
        return (self.ite % self.param.nhis == 0) or self.finished
