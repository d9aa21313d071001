"""A user-style driver in the reference's plugin format (drivers/template.py): v1 signature."""


class Driver:
    def __init__(self):
        self.calls = 0

    def process_lidar(self, ranges):
        self.calls += 1
        if self.calls % 50 == 0:
            raise RuntimeError("flaky driver")          # the simulator prints the error and keeps the last controls
        best = int(ranges[11:79].argmax()) + 11
        steering = (best - 45) * (3.141592653589793 * 2 / 90)
        return 0.8, max(-0.6, min(0.6, steering))
