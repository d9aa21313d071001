"""A v2 driver: process_lidar(ranges, state) with the VehicleStateSnapshot fields (ft_grandprix/vehicle.py:3-12)."""


class Driver:
    def process_lidar(self, ranges, state):
        assert hasattr(state, "laps") and hasattr(state, "yaw") and len(state.velocity) == 3
        best = int(ranges[11:79].argmax()) + 11
        return 1.0 if state.lap_completion >= 0 else 0.5, (best - 45) * 0.07
