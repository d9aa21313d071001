"""Headless race runner: the part of the reference's GUI that is race output.

    python -m ft_grandprix_b200.run --cars template/cars/cars.json [--track track] [--lap-target 10]
                                    [--max-seconds 120] [--worlds 1] [--report-every 5]

Reads a cars JSON list of {driver, name, primary, secondary, icon} exactly as the reference does
(template/cars/cars.json; ft_grandprix/custom.py:1095-1121): `driver` is a dotted module path or `file://path/to/x.py`,
loaded with importlib; an import failure gives an inert LobotomyDriver (custom.py:1106-1109).  The bundled drivers
(ft_grandprix.nidc / .fast / .lobotomy) run on the device; any other Driver object runs unchanged on the host through
Fleet.drive_host (ranges copied to the host each tick -- the slow path the plugin API needs).  Every `--worlds` world holds
one car per JSON entry on the reference start grid (car i at path[(i+5)*2], custom.py:1232-1245); the cars see each
other's lidar cylinders; car-car contacts are not generated.  Prints, per car, what the dashboard
shows (custom.py:335-361): position by absolute completion, laps, completion %, lap times, and the finish line text.
"""
import argparse
import importlib
import importlib.util
import json
import os
import sys

import numpy as np

from .fleet import DRIVER_KINDS, LAP, TIMESTEP, Fleet
from .track import Track


def ordinal(n):
    """'1st', '2nd', '3rd', '4th', '11th', ... as the dashboard prints positions (behaviour of custom.py:47-55)."""
    n = int(n)
    suffix = "th" if 10 <= n % 100 <= 20 else {1: "st", 2: "nd", 3: "rd"}.get(n % 10, "th")
    return f"{n}{suffix}"


class LobotomyDriver:                                # ft_grandprix/lobotomy.py:1-3
    def process_lidar(self, ranges):
        return 0.0, 0.0


def load_driver(spec):
    """(device kind or None, host driver object or None, module path) for one cars.json entry (custom.py:1097-1109)."""
    if spec.startswith("file://"):
        path = spec[7:-3].replace("/", ".")
    elif "//" not in spec:
        path = spec
    else:
        print("Unsupported schema: supported (file://)")
        return DRIVER_KINDS["lobotomy"], None, spec
    print(f"Loading driver from python module path '{path}'")
    if path in DRIVER_KINDS:
        return DRIVER_KINDS[path], None, path
    try:
        if spec.startswith("file://") and os.path.exists(spec[7:]):
            s = importlib.util.spec_from_file_location(os.path.basename(spec[7:-3]), spec[7:])
            module = importlib.util.module_from_spec(s)
            s.loader.exec_module(module)
        else:
            module = importlib.import_module(path)
        return None, module.Driver(), path
    except Exception:                                # custom.py:1106-1109: any failure gives an inert driver
        return DRIVER_KINDS["lobotomy"], None, path


def dashboard(fleet, names, world=0):
    """The per-car lines of the reference's vehicle list for one world, sorted by position."""
    cpw = fleet.cars_per_world
    lap = fleet.lap[world * cpw:(world + 1) * cpw].cpu().numpy()
    times = fleet.times[world * cpw:(world + 1) * cpw].cpu().numpy()
    comp = np.where(lap[:, LAP["good_start"]] != 0, lap[:, LAP["completion"]], lap[:, LAP["completion"]] - 100)
    absolute = lap[:, LAP["laps"]] * 100 + comp
    order = sorted(range(cpw), key=lambda i: absolute[i], reverse=True)      # custom.py:335
    lines = []
    for pos, i in enumerate(order):
        t = [float(x) * TIMESTEP for x in times[i, :lap[i, LAP["ntimes"]]]]
        ts = "[" + ", ".join(f"{x:.2f}" for x in t) + "]"
        head = f"{ordinal(pos + 1):>4}  Car #{i} - {names[i]}"
        if lap[i, LAP["finished"]]:
            lines.append(f"{head}: Car finished {lap[i, LAP['laps']]} laps in {sum(t):.2f} seconds!  Lap Times: {ts}")
        else:
            lines.append(f"{head}: Laps: {lap[i, LAP['laps']]}  Completion: {comp[i]}%  Lap Times: {ts}")
    return lines


def run(cars, track="track", lap_target=10, max_seconds=120.0, worlds=1, report_every=5.0, device=0, out=print):
    specs = [load_driver(c["driver"]) for c in cars]
    names = [c.get("name", f"car {i}") for i, c in enumerate(cars)]
    cpw = len(cars)
    t = Track.bundled(track) if isinstance(track, str) else track
    fleet = Fleet(t, cpw * worlds, cars_per_world=cpw, device=device, lap_target=lap_target)
    fleet.reset_grid()
    host = [(i, d) for i, (k, d, _) in enumerate(specs) if d is not None]
    fleet.set_driver_kinds([DRIVER_KINDS["lobotomy"] if k is None else k for k, _, _ in specs] * worlds)
    max_ticks = int(round(max_seconds / TIMESTEP))
    every = max(1, int(round(report_every / TIMESTEP)))
    tick = 0
    while tick < max_ticks:
        n = min(every, max_ticks - tick)
        if not host:
            fleet.tick(n)                            # all drivers on the device: whole iterations without leaving it
        else:
            import torch
            drivers = [None] * fleet.ncars
            for w in range(worlds):
                for i, d in host:
                    drivers[w * cpw + i] = d
            hidx = torch.tensor([i for i, d in enumerate(drivers) if d is not None], device=fleet.device)
            for _ in range(n):                       # custom.py:1337-1426 with the plugin call on the host
                fleet.lap_update()
                with torch.cuda.stream(fleet.stream):
                    keep = fleet.ctrl[hidx].clone()  # a host driver that raises leaves last tick's controls in place
                    fleet.drive()
                    fleet.ctrl[hidx] = keep
                _drive_host_subset(fleet, drivers)
                fleet.lidar()
                fleet.step(1)
        tick += n
        fleet.sync()
        out(f"t = {tick * TIMESTEP:.1f} s")
        for line in dashboard(fleet, names):
            out("  " + line)
        if bool((fleet.lap[:, LAP["finished"]] != 0).all()):
            break
    return fleet


def _drive_host_subset(fleet, drivers):
    """Fleet.drive_host for the cars that have a host driver (None entries keep the device driver's ctrl); finished
    cars are shadowed: their driver is replaced by the inert one (custom.py:1437)."""
    import inspect
    fleet.sync()
    ranges = fleet.ranges.cpu().numpy().astype(np.float64)
    ctrl = fleet.ctrl.cpu().numpy()
    finished = fleet.lap[:, LAP["finished"]].cpu().numpy()
    snaps = None
    for i, d in enumerate(drivers):
        if d is None:
            continue
        if finished[i]:
            ctrl[i] = (0.0, 0.0)
            continue
        try:
            if len(inspect.signature(d.process_lidar).parameters) >= 2:
                if snaps is None:
                    snaps = fleet.snapshots()
                sp, st = d.process_lidar(ranges[i].copy(), snaps[i])
            else:
                sp, st = d.process_lidar(ranges[i].copy())
            ctrl[i] = (float(sp), float(st))
        except Exception as e:                       # custom.py:1409-1411
            print(f"Error in vehicle `{i}`: `{e}`")
    import torch
    with torch.cuda.stream(fleet.stream):
        fleet.ctrl.copy_(torch.from_numpy(ctrl).to(fleet.device))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--cars", required=True, help="cars JSON list (template/cars/cars.json format)")
    ap.add_argument("--track", default="track", choices=["track", "circle", "small-circle", "inkscape"])
    ap.add_argument("--lap-target", type=int, default=10)
    ap.add_argument("--max-seconds", type=float, default=120.0)
    ap.add_argument("--worlds", type=int, default=1, help="identical copies of the race run side by side")
    ap.add_argument("--report-every", type=float, default=5.0, help="simulated seconds between dashboard prints")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    cars = json.load(open(a.cars))
    run(cars, a.track, a.lap_target, a.max_seconds, a.worlds, a.report_every, a.device)


if __name__ == "__main__":
    main()
