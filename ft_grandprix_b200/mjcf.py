"""MJCF / `rendered/` emitter: writes what the reference's stage() leaves on disk before it calls
MjModel.from_xml_path (ft_grandprix/custom.py:1154-1179):

    rendered/chunks/XXXxYYY.png + metadata.json   ft_grandprix/chunk.py:45-80
    rendered/car.xml, rendered/car.json           ft_grandprix/map.py:10-72 expanding template/mushr.em.xml
                                                  (or template/car.em.xml in tricycle mode)
    rendered/meshes/*.stl, rendered/icons/*.png   map.py:27-40 (copies of template/meshes, template/icons)

so that anyone with a MuJoCo install can compile the very world this package simulates and capture goldens from it
(tests/golden/make_mujoco_golden.py).  No empy: the model is described below as plain Python data (bodies, joints,
geoms ...) and serialised by a 20-line XML writer.  Numbers are produced by the same float expressions the templates
evaluate and printed with str(), so every attribute string equals the empy expansion's; tests/test_mjcf.py checks
the element tree against goldens expanded from the reference's own template files.
"""
import json
import os
import struct
from math import ceil, cos, radians, sin

import numpy as np

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

# CSS-like names accepted for "primary" / "secondary" in a cars JSON (ft_grandprix/colors.py is a 140-entry table; the
# entries the bundled car lists use are kept here, anything else must be given as [r, g, b] or "rgb(r,g,b)")
COLOR_NAMES = {
    "red": [255, 0, 0], "pink": [255, 192, 203], "orange": [255, 165, 0], "darkorange": [255, 140, 0],
    "blue": [0, 0, 255], "green": [0, 128, 0], "white": [255, 255, 255], "black": [0, 0, 0], "yellow": [255, 255, 0],
    "purple": [128, 0, 128], "gray": [128, 128, 128], "grey": [128, 128, 128], "cyan": [0, 255, 255],
    "magenta": [255, 0, 255], "brown": [165, 42, 42], "gold": [255, 215, 0], "error": [255, 200, 200],
    "lightblue": [173, 216, 230], "maroon": [128, 0, 0],
}


def resolve_color(color):
    """colors.resolve_color (ft_grandprix/colors.py:140-145) without the random branch."""
    if isinstance(color, (list, tuple)):
        return list(color)
    if color.startswith("rgb"):
        return [int(x) for x in color[4:-1].split(",")]
    return list(COLOR_NAMES[color])


# ---------------------------------------------------------------------------------------------- tiny XML writer
class E:
    """element: tag, attributes in insertion order, children"""
    def __init__(self, tag, attrs=None, children=None, **kw):
        self.tag, self.attrs, self.children = tag, dict(attrs or {}), list(children or [])
        self.attrs.update(kw)

    def add(self, *els):
        self.children.extend(els)
        return self

    def write(self, out, depth=0):
        pad = "  " * depth
        a = "".join(f' {k}="{v}"' for k, v in self.attrs.items())
        if not self.children:
            out.append(f"{pad}<{self.tag}{a}/>\n")
            return
        out.append(f"{pad}<{self.tag}{a}>\n")
        for c in self.children:
            c.write(out, depth + 1)
        out.append(f"{pad}</{self.tag}>\n")


def _v(*xs):
    return " ".join(str(x) for x in xs)


def _chunk_name(x, y):
    return f"{x:03}x{y:03}"


# ---------------------------------------------------------------------------------------------- mushr (default) model
def mushr_model(cars, metadata, rangefinders=90, map_color=(1, 0, 0)):
    """The element tree of rendered/car.xml for template/mushr.em.xml (line numbers below are that file's)."""
    scale = metadata["scale"]
    border_height = 0.2                                                      # :16
    map_w = map_h = 20 * scale                                               # :17
    size_x = map_w / metadata["horizontal_chunks"]                           # :19
    size_y = map_h / metadata["vertical_chunks"]                             # :20
    inter_ray_angle = 360 / rangefinders                                     # :21
    affordance, ms, wr = 0.1, 0.5, 0.03                                      # :22-24
    pr, pg, pb = map_color
    chunks = [tuple(c) for c in metadata["chunks"]]
    n = len(cars)

    root = E("mujoco")
    root.add(E("compiler", angle="radian"), E("option", timestep="0.004"),
             E("visual").add(E("headlight", ambient="0.5 0.5 0.5"), E("quality", shadowsize="4096"), E("map", znear="0.001")))
    asset = E("asset")
    asset.add(E("material", name="transparent", rgba="0 0 0 0"),
              E("mesh", scale=_v(ms, ms, ms), name="buddy_mushr_base_nano", file="meshes/simple_base_nano.stl"),
              E("mesh", scale=_v(ms * 1.3, ms * 1.3, ms * 1.3), name="buddy_mushr_wheel", file="meshes/mushr_wheel.stl"))
    for i, car in enumerate(cars):
        r1, g1, b1 = [x / 255 for x in car["primary"]]
        r2, g2, b2 = [x / 255 for x in car["secondary"]]
        asset.add(E("texture", type="2d", name=f"car #{i} icon", file=f"icons/{car['icon']}"),
                  E("material", name=f"car #{i} icon", rgba="0.8 0.8 0.8 1.0", texture=f"car #{i} icon", reflectance="1"),
                  E("material", name=f"car #{i} primary", rgba=_v(r1, g1, b1, "1.0")),
                  E("material", name=f"car #{i} secondary", rgba=_v(r2, g2, b2, "0.5")),
                  E("material", name=f"car #{i} body", rgba="0.1 0.1 0.1 1.0"),
                  E("material", name=f"car #{i} wheel", rgba="0.1 0.1 0.1 1.0"))
    for x, y in chunks:
        asset.add(E("hfield", name=f"map-{_chunk_name(x, y)}", file=f"chunks/{_chunk_name(x, y)}.png",
                    size=_v(size_x / 2, size_y / 2, border_height + affordance, "0.0001")))
    asset.add(E("texture", type="2d", name="grid", builtin="checker", rgb1=_v(pr, pg, pb), rgb2="1 1 1", width="512", height="512"),
              E("material", name="plane", reflectance="0.1", rgba=_v(pr / 3, pg / 3, pb / 3, 1)),
              E("material", name="grid", texture="grid", texuniform="true", reflectance="0"))
    root.add(asset)

    def dflt(cls, el):
        return E("default", {"class": cls}).add(el)
    root.add(E("default").add(
        dflt("buddy_suspension", E("joint", type="slide", axis="0 0 1", frictionloss="0.001", stiffness="500.0", springref="-0.015",
                                   damping="12.5", armature="0.01", range="-0.03 0")),
        dflt("buddy_softener", E("geom", mesh="buddy_mushr_wheel", mass="0.00001", rgba="1 1 1 0", fitscale="2.0", contype="0", conaffinity="4")),
        dflt("buddy_wheel", E("geom", type="ellipsoid", size=_v(wr, "0.01", "0.03"), friction="0.3 0.005 0.0001", condim="3", contype="1",
                              conaffinity="0", mass="0.498952", solimp="0 0.95 0.001 0.5 2", solref="0.02 1", margin="0")),
        dflt("buddy_wheel_site", E("site", type="cylinder", size=_v(wr, "0.01"), euler=_v(-radians(90), 0, 0))),
        dflt("buddy_steering", E("joint", type="hinge", axis="0 0 1", limited="true", frictionloss="0.01", damping="0.1",
                                 armature="0.0002", range="-1 1")),
        dflt("buddy_throttle", E("joint", type="hinge", axis="0 1 0", frictionloss="0.001", damping="0.01", armature="0.01", limited="false")),
        dflt("wheel", E("geom", size="0.03 0.01 0", type="cylinder")),
        dflt("decor", E("site", type="box"))))

    world = E("worldbody")
    for x, y in chunks:
        world.add(E("geom", conaffinity="1", contype="4", material="grid", type="hfield", hfield=f"map-{_chunk_name(x, y)}",
                    pos=_v(size_x * x, -size_y * y, -affordance)))
    world.add(E("geom", friction=".5 0.005 0.0001", conaffinity="3", contype="3", name="plane", size="300 300 0.1", pos="0 0 0.01",
                type="plane", material="plane"))
    rx, ry, rz = -0.0525, 0.000, 0.065                                       # rangefinder origin :101
    lr, lh = 0.030, 0.015                                                    # lidar model radius / height :103-104
    for i in range(n):
        body = E("body", name=f"car #{i}", pos="0.0 2.0 0.0", euler="0 0 0.0")
        body.add(E("freejoint", name=f"car #{i}"),
                 E("light", name=f"top light #{i}", pos=_v(rx, ry, rz + 0.1), dir="0 0 -1", diffuse="0.5 0.5 0.5", mode="trackcom"),
                 E("light", name=f"front light #{i}", pos=_v(rx, ry, rz + 0.1), dir="0.894427 0 -0.447214", diffuse="1 1 1"),
                 E("geom", name=f"car #{i} lidar", type="cylinder", size=_v(lr, lh, "0.01"), pos=_v(rx, ry, rz - lh / 2), material=f"car #{i} body"),
                 E("site", name=f"buddy_imu #{i}", pos=_v(rx, ry, rz - lh / 2), rgba="0 0 0 0"),
                 E("site", material=f"car #{i} icon", type="ellipsoid", euler=_v(0, 0, radians(-90)), size=_v(lr * 0.9, lr * 0.9, "0.0075"),
                   pos=_v(rx, ry, rz + lh / 2 - 0.0005)),
                 E("site", material=f"car #{i} secondary", type="ellipsoid", euler=_v(0, 0, radians(-90)), size=_v(lr * 1.1, lr * 1.1, "0.0100"),
                   pos=_v(rx, ry, rz)))
        for j in range(rangefinders):
            theta = -radians(inter_ray_angle * j - 90)
            body.add(E("site", name=f"rangefinder #{i}.#{j}", pos=_v(rx + lr * sin(theta), lr * cos(theta), rz),
                       euler=_v(radians(90), radians(inter_ray_angle * j - 90), 0), rgba="0 0 0 0"))
        body.add(E("camera", name=f"buddy_third_person #{i}", mode="fixed", pos="-1 0 1", xyaxes="0 -1 0 0.707 0 0.707"),
                 E("geom", name=f"chasis #{i}", pos=_v(0, 0, ms * 0.094655), type="mesh", mass="3.542137", mesh="buddy_mushr_base_nano",
                   material=f"car #{i} primary"),
                 E("body", name=f"buddy_steering_wheel #{i}", pos="0.1385 0 0.0488").add(
                     E("joint", {"class": "buddy_steering"}, name=f"buddy_steering_wheel #{i}"),
                     E("geom", {"class": "buddy_wheel"}, contype="0", conaffinity="0", mass="0.01", rgba="0 0 0 0")))
        for tag, px, py, steers in (("fl", 0.1385, 0.115, True), ("fr", 0.1385, -0.115, True), ("bl", -0.158, 0.115, False), ("br", -0.158, -0.115, False)):
            wheel = E("body", name=f"buddy_wheel_{tag} #{i}", pos=_v(ms * px, ms * py, ms * 0.0488))
            wheel.add(E("joint", {"class": "buddy_suspension"}, name=f"buddy_wheel_{tag}_suspension #{i}"),
                      E("body").add(E("joint", type="ball", frictionloss="0.25"),
                                    E("geom", name=f"{tag} softener #{i}", **{"class": "buddy_softener"})))
            if steers:
                wheel.add(E("joint", {"class": "buddy_steering"}, name=f"buddy_wheel_{tag}_steering #{i}"))
            wheel.add(E("joint", {"class": "buddy_throttle"}, name=f"buddy_wheel_{tag}_throttle #{i}"),
                      E("geom", {"class": "buddy_wheel"}, material=f"car #{i} secondary", name=f"buddy_wheel_{tag}_throttle #{i}", group="3"),
                      E("site", {"class": "buddy_wheel_site"}, material=f"car #{i} wheel"),
                      E("site", type="box", material=f"car #{i} secondary", size=_v(wr * 0.9, "0.011", "0.007")),
                      E("site", type="box", material=f"car #{i} secondary", size=_v("0.007", "0.011", wr * 0.9)))
            body.add(wheel)
        world.add(body)
    root.add(world)

    act, eq, ten, sen = E("actuator"), E("equality"), E("tendon"), E("sensor")
    for i in range(n):
        act.add(E("position", {"class": "buddy_steering"}, kp="20.0", name=f"turn #{i}", joint=f"buddy_steering_wheel #{i}"),
                E("velocity", kv="100", gear="0.04", forcelimited="true", forcerange="-500 500", name=f"forward #{i}", tendon=f"buddy_throttle #{i}"))
        eq.add(E("joint", joint1=f"buddy_wheel_fl_steering #{i}", joint2=f"buddy_steering_wheel #{i}", polycoef="0 1 0.375 0.140625 -0.0722656"),
               E("joint", joint1=f"buddy_wheel_fr_steering #{i}", joint2=f"buddy_steering_wheel #{i}", polycoef="0 1 -0.375 0.140625 0.0722656"))
        ten.add(E("fixed", name=f"buddy_throttle #{i}").add(
            *[E("joint", joint=f"buddy_wheel_{tag}_throttle #{i}", coef="0.25") for tag in ("fl", "fr", "bl", "br")]))
        for j in range(rangefinders):                                        # all rangefinders of all cars first :204-206
            sen.add(E("rangefinder", name=f"rangefinder #{i}.#{j}", site=f"rangefinder #{i}.#{j}"))
    for i in range(n):
        sen.add(E("gyro", name=f"car #{i} gyro", site=f"buddy_imu #{i}"),
                E("accelerometer", name=f"car #{i} accelerometer", site=f"buddy_imu #{i}"))
    for i in range(n):
        sen.add(E("accelerometer", name=f"buddy_accelerometer #{i}", site=f"buddy_imu #{i}"),
                E("gyro", name=f"buddy_gyro #{i}", site=f"buddy_imu #{i}"),
                E("velocimeter", name=f"buddy_velocimeter #{i}", site=f"buddy_imu #{i}"))
    root.add(act, eq, ten, sen)
    return root


# ---------------------------------------------------------------------------------------------- tricycle model
def tricycle_model(cars, metadata, rangefinders=90, map_color=(1, 0, 0)):
    """The element tree of rendered/car.xml for template/car.em.xml (option tricycle_mode, custom.py:1163-1170)."""
    size_x = metadata["scale"] * 20 / metadata["horizontal_chunks"]          # car.em.xml:3-4
    size_y = metadata["scale"] * 20 / metadata["vertical_chunks"]
    inter_ray_angle = 360 / rangefinders
    affordance = 0.1
    pr, pg, pb = map_color
    chunks = [tuple(c) for c in metadata["chunks"]]
    n = len(cars)
    root = E("mujoco", model="MuJoCo Model")
    root.add(E("compiler", autolimits="true", texturedir="icons/"), E("option", timestep="0.0075"),
             E("visual").add(E("rgba", rangefinder="1.0 1.0 0.0 0.075"), E("headlight", ambient="0.5 0.5 0.5"), E("quality", shadowsize="4096")),
             E("statistic", meansize="0.509902", extent="1.6099", center="0 0 0.704951"))
    root.add(E("default", {"class": "main"}).add(
        E("default", {"class": "suspension"}).add(E("joint", type="slide", stiffness="10", damping="0.5", armature="0.01", range="-0.01 -0.001")),
        E("joint", damping="0.03"),
        E("default", {"class": "wheel"}).add(E("geom", solimp="0 0.95 0.001 0.5 2", size="0.03 0.01 0", type="cylinder")),
        E("default", {"class": "decor"}).add(E("site", type="box")),
        E("default", {"class": "tester"})))
    asset = E("asset").add(E("material", name="transparent", rgba="0 0 0 0"))
    for i, car in enumerate(cars):
        r1, g1, b1 = [x / 255 for x in car["primary"]]
        r2, g2, b2 = [x / 255 for x in car["secondary"]]
        asset.add(E("texture", type="2d", name=f"car #{i} icon", file=f"{car['icon']}"),
                  E("material", name=f"car #{i} icon", rgba="0.8 0.8 0.8 1.0", texture=f"car #{i} icon", reflectance="1"),
                  E("material", name=f"car #{i} primary", rgba=_v(r1, g1, b1, "1.0")),
                  E("material", name=f"car #{i} secondary", rgba=_v(r2, g2, b2, "1.0")),
                  E("material", name=f"car #{i} body", rgba="0.1 0.1 0.1 1.0"),
                  E("material", name=f"car #{i} wheel", rgba="0.1 0.1 0.1 1.0"))
    asset.add(E("texture", type="2d", name="grid", builtin="checker", rgb1=_v(pr, pg, pb), rgb2="1 1 1", width="512", height="512"),
              E("material", name="plane", reflectance="0.1", rgba=_v(pr / 3, pg / 3, pb / 3, 1)),
              E("material", name="grid", texture="grid", texuniform="true", reflectance="0"),
              E("mesh", name="chasis", vertex="9 2 0 -10 10 10 9 -2 0 10 3 -10 10 -3 -10 -8 10 -10 -10 -10 10 -8 -10 -10 -5 0 20",
                scale="0.01 0.006 0.0015"))
    for x, y in chunks:
        asset.add(E("hfield", name=f"map-{_chunk_name(x, y)}", file=f"chunks/{_chunk_name(x, y)}.png",
                    size=_v(size_x / 2, size_y / 2, 0.15 + affordance, "0.0001")))
    root.add(asset)
    world = E("worldbody")
    for x, y in chunks:
        world.add(E("geom", conaffinity="5", contype="4", material="grid", type="hfield", hfield=f"map-{_chunk_name(x, y)}",
                    pos=_v(size_x * x, -size_y * y, -affordance)))
    world.add(E("geom", conaffinity="3", contype="1", name="plane", material="plane", size="300 300 0.1", pos="0 0 0.01", type="plane", rgba="0.3 0 0 1"),
              E("light", pos="4.5 -3.0 3", dir="0 0 -1", diffuse="0.5 0.5 0.5"))
    rx, ry, rz = -0.0525, 0.000, 0.030
    lr, lh = 0.030, 0.015
    mw = 0.5
    for i, car in enumerate(cars):
        body = E("body", name=f"car #{i}", pos=_v(car["x"], car["y"], car["z"]))
        body.add(E("joint", name=f"car #{i}", type="free", damping="0"),
                 E("light", name=f"top light #{i}", pos="0 0 2", dir="0 0 -1", diffuse="0.4 0.4 0.4", mode="trackcom"),
                 E("light", name=f"front light #{i}", pos="0.1 0 0.02", dir="0.894427 0 -0.447214", diffuse="1 1 1"),
                 E("geom", name=f"chasis #{i}", type="mesh", mesh="chasis", material=f"car #{i} primary"),
                 E("site", name=f"car #{i} accelerometer", euler="0 0 -90", rgba="0 0 0 0"),
                 E("site", name=f"car #{i} gyroscope", euler="0 0 -90", rgba="0 0 0 0"),
                 E("geom", density="2000", name=f"car #{i} lidar", type="cylinder", size=_v(lr, lh, "0.01"), pos=_v(rx, ry, rz - lh / 2),
                   material=f"car #{i} body"),
                 E("site", material=f"car #{i} icon", type="ellipsoid", euler="0 0 -90", size=_v(lr * 0.9, lr * 0.9, "0.0075"),
                   pos=_v(rx, ry, rz + lh / 2 - 0.0005)),
                 E("site", material=f"car #{i} secondary", type="ellipsoid", euler="0 0 -90", size=_v(lr * 1.1, lr * 1.1, "0.0100"), pos=_v(rx, ry, rz)))
        for j in range(rangefinders):
            theta = -radians(inter_ray_angle * j - 90)
            body.add(E("site", name=f"rangefinder #{i}.#{j}", pos=_v(rx + lr * sin(theta), lr * cos(theta), rz),
                       euler=_v(90, inter_ray_angle * j - 90, 0), rgba="0 0 0 0"))
        softener = lambda name: E("geom", size="0.035", rgba="0.7 0.7 0.7 0.0", contype="0", conaffinity="4", mass="0.00000001", name=name)
        body.add(E("body", pos="0.090 0 -0.01").add(E("joint", {"class": "suspension"}), E("joint", type="ball"), softener(f"front softener #{i}")),
                 E("geom", name=f"front wheel #{i}", mass=_v(mw / 3), material=f"car #{i} wheel", size="0.015", pos="0.08 0 -0.015", condim="1", priority="1"))
        for side, py, named in (("left", "0.06", True), ("right", "-0.06", False)):
            attrs = {"name": f"{side} wheel #{i}"} if named else {}
            attrs.update(pos=f"-0.07 {py} 0", quat="0.707107 -0.707107 0 0")
            body.add(E("body", attrs).add(
                E("joint", {"class": "suspension"}, axis="0 -1 0"),
                E("body").add(E("joint", type="ball"), softener(f"{side} softener #{i}")),
                E("joint", name=f"{side} #{i}", pos="0 0 0", axis="0 0 1"),
                E("geom", {"name": f"{side} wheel #{i}", "material": f"car #{i} wheel", "mass": _v(mw / 3), "class": "wheel"}),
                E("site", {"class": "decor"}, material=f"car #{i} secondary", pos="0 0 0", size="0.006 0.025 0.012"),
                E("site", {"class": "decor"}, material=f"car #{i} secondary", pos="0 0 0", size="0.025 0.006 0.012")))
        world.add(body)
    root.add(world)
    ten, act, sen = E("tendon"), E("actuator"), E("sensor")
    for i in range(n):
        ten.add(E("fixed", name=f"forward #{i}").add(E("joint", joint=f"left #{i}", coef="0.5"), E("joint", joint=f"right #{i}", coef="0.5")),
                E("fixed", name=f"turn #{i}").add(E("joint", joint=f"left #{i}", coef="-0.5"), E("joint", joint=f"right #{i}", coef="0.5")))
        act.add(E("motor", name=f"forward #{i}", tendon=f"forward #{i}", ctrlrange="-4 4"),
                E("motor", name=f"turn #{i}", tendon=f"turn #{i}", ctrlrange="-1 1"))
        sen.add(E("jointactuatorfrc", joint=f"right #{i}", name=f"right #{i}"), E("jointactuatorfrc", joint=f"left #{i}", name=f"left #{i}"))
        for j in range(rangefinders):
            sen.add(E("rangefinder", name=f"rangefinder #{i}.#{j}", site=f"rangefinder #{i}.#{j}"))
    for i in range(n):
        sen.add(E("gyro", name=f"car #{i} gyro", site=f"car #{i} accelerometer"),
                E("accelerometer", name=f"car #{i} accelerometer", site=f"car #{i} gyroscope"))
    root.add(ten, act, sen)
    return root


# ---------------------------------------------------------------------------------------------- rendered/ directory
def prepare_cars(cars):
    """What produce_mjcf does to the car list before templating (map.py:33-45): colours resolved, start x/y/z added."""
    out = []
    for index, car in enumerate(cars):
        car = dict(car)
        for color in ("primary", "secondary"):
            car[color] = resolve_color(car[color])
        car["x"] = 4.5 + 5.5 + 0.1 * (index % 3)
        car["y"] = -8.5 + 0.0 + 0.1 * (index % 3)
        car["z"] = 0.1
        out.append(car)
    return out


def write_chunks(wall, output_dir, name="track", scale=2.0, chunk_width=20, chunk_height=20):
    """chunk() (chunk.py:38-80) from the wall mask (pixel is wall iff R+G+B == 765): one RGB PNG per non-empty chunk
    (walls 255, everything else 0) plus metadata.json.  Returns the metadata dict."""
    from PIL import Image
    wall = np.asarray(wall).astype(bool)
    h, w = wall.shape
    rgb = np.repeat((wall.astype(np.uint8) * 255)[:, :, None], 3, axis=2)
    os.makedirs(output_dir, exist_ok=True)
    hc, vc = ceil(w / chunk_width), ceil(h / chunk_height)
    chunks = []
    for i in range(hc):
        for j in range(vc):
            tile = rgb[j * chunk_height:min((j + 1) * chunk_height, h), i * chunk_width:min((i + 1) * chunk_width, w)]
            if tile.sum() > 0:
                Image.fromarray(np.ascontiguousarray(tile)).save(os.path.join(output_dir, f"{_chunk_name(i, j)}.png"))
                chunks.append([i, j])
    metadata = {"original_width": w, "original_height": h, "chunk_width": chunk_width, "chunk_height": chunk_height,
                "horizontal_chunks": hc, "vertical_chunks": vc, "chunks": chunks, "width": w, "height": h,
                "name": name, "scale": scale}
    with open(os.path.join(output_dir, "metadata.json"), "w") as f:
        json.dump(metadata, f)
    return metadata


def write_stl(path, triangles, normals=None):
    """binary STL: 80-byte header, uint32 count, per triangle normal + 3 vertices (float32) + uint16 0"""
    tri = np.asarray(triangles, dtype="<f4").reshape(-1, 3, 3)
    if normals is None:
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    else:
        nrm = np.asarray(normals, dtype="<f4").reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(b"ft_grandprix_b200 mesh".ljust(80, b" "))
        f.write(struct.pack("<I", len(tri)))
        for t, nn in zip(tri, nrm.astype("<f4")):
            f.write(nn.tobytes()); f.write(t.tobytes()); f.write(b"\0\0")


def produce_mjcf(cars, metadata, output_dir="rendered", rangefinders=90, map_color=(1, 0, 0), tricycle=False,
                 template_dir=None):
    """produce_mjcf (map.py:10-72): writes car.xml, car.json, meshes/ and icons/ into output_dir.  With template_dir (a
    checkout of the reference's template/) meshes and icons are copied from there like the reference does; otherwise
    the bundled mesh data is written as binary STL and plain white icons are generated."""
    os.makedirs(output_dir, exist_ok=True)
    cars = prepare_cars(cars)
    model = (tricycle_model if tricycle else mushr_model)(cars, metadata, rangefinders, map_color)
    out = []
    model.write(out)
    with open(os.path.join(output_dir, "car.xml"), "w") as f:
        f.write("".join(out))
    with open(os.path.join(output_dir, "car.json"), "w") as f:
        json.dump({"cars": cars, "rangefinders": rangefinders}, f)
    mesh_dir, icon_dir = os.path.join(output_dir, "meshes"), os.path.join(output_dir, "icons")
    if template_dir is not None:
        import shutil
        for d in (mesh_dir, icon_dir):
            if os.path.isdir(d):
                shutil.rmtree(d)
        shutil.copytree(os.path.join(template_dir, "meshes"), mesh_dir)
        shutil.copytree(os.path.join(template_dir, "icons"), icon_dir)
    else:
        from PIL import Image
        os.makedirs(mesh_dir, exist_ok=True); os.makedirs(icon_dir, exist_ok=True)
        z = np.load(os.path.join(ASSETS, "meshes.npz"))
        for name in ("simple_base_nano", "mushr_wheel"):
            write_stl(os.path.join(mesh_dir, name + ".stl"), z[name + "__tri"], z[name + "__nrm"])
        for car in cars:
            if car.get("icon"):
                Image.new("RGB", (64, 64), (255, 255, 255)).save(os.path.join(icon_dir, car["icon"]))
    return os.path.join(output_dir, "car.xml")


def render_world(track, cars, output_dir="rendered", tricycle=False, map_color=(1, 0, 0), template_dir=None):
    """stage() up to the MuJoCo compile (custom.py:1154-1170): chunk() then produce_mjcf() for a Track of this package."""
    metadata = write_chunks(track.wall, os.path.join(output_dir, "chunks"), name=track.name, scale=track.scale,
                            chunk_width=track.chunk_px, chunk_height=track.chunk_px)
    return produce_mjcf(cars, metadata, output_dir, rangefinders=90, map_color=map_color, tricycle=tricycle,
                        template_dir=template_dir)


if __name__ == "__main__":
    import argparse
    from .track import Track
    ap = argparse.ArgumentParser(description="write rendered/ (chunks, car.xml, car.json, meshes, icons) for a bundled track")
    ap.add_argument("--track", default="track")
    ap.add_argument("--cars", default=None, help="cars JSON (list of {driver, name, primary, secondary, icon}); default: one nidc car")
    ap.add_argument("--output-dir", default="rendered")
    ap.add_argument("--tricycle", action="store_true")
    ap.add_argument("--template-dir", default=None)
    a = ap.parse_args()
    cars = json.load(open(a.cars)) if a.cars else [{"driver": "ft_grandprix.nidc", "name": "car", "primary": "red", "secondary": "pink", "icon": "white.png"}]
    print(render_world(Track.bundled(a.track), cars, a.output_dir, tricycle=a.tricycle, template_dir=a.template_dir))
