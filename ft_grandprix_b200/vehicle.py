"""Per-car snapshot handed to v2 drivers: same fields, same meaning as the reference's
VehicleStateSnapshot (ft_grandprix/vehicle.py:3-12, filled at ft_grandprix/custom.py:149-160)."""
from dataclasses import dataclass


@dataclass
class VehicleStateSnapshot:
    laps: int
    velocity: list          # world-frame linear velocity of the free joint, qvel[0:3]
    yaw: float
    pitch: float
    roll: float
    lap_completion: int     # negative while driving a lap backwards (custom.py:132-140)
    absolute_completion: int
    time: float
