#!/usr/bin/env python
"""Golden vectors for the MJCF / rendered/ emitter (ft_grandprix_b200/mjcf.py).  BUILD CONTAINER ONLY.

The reference renders `rendered/car.xml` by expanding template/mushr.em.xml (or car.em.xml) with empy
(ft_grandprix/map.py:55-65).  empy is not installable here, so this script carries a minimal interpreter for the empy
subset the two templates use -- `@{ statements }`, `@[for ...] ... @[end for]`, `@(expr)`, `@name`, `@name['k']`,
`@<newline>` / `@<space>` -- and runs it on the REFERENCE'S OWN TEMPLATE FILES, read from /root/reference at
generation time, with the locals produce_mjcf() passes (map.py:58-64).  The expansion is parsed as XML and written in a
canonical form (tags, sorted attributes, nesting; comments and whitespace dropped) to

    tests/golden/mjcf_<template>_<cars>_<track>.txt.gz

tests/test_mjcf.py expects the emitter's car.xml, canonicalised the same way, to be IDENTICAL.  Also written:
rendered/chunks/metadata.json and two chunk PNGs per track as produced by ft_grandprix.chunk.chunk() itself
(tests/golden/chunks_rendered.json, sha256 of every file), and rendered/car.json.

    python tests/golden/make_mjcf_golden.py
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile
import xml.etree.ElementTree as ET

REF = os.environ.get("FTGP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------- mini empy
def _match(src, i, open_c, close_c):
    """index just past the bracket that closes src[i] == open_c (string-aware)."""
    depth, k, n = 0, i, len(src)
    while k < n:
        c = src[k]
        if c in "'\"":
            q = c
            k += 1
            while k < n and src[k] != q:
                k += 2 if src[k] == "\\" else 1
        elif c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return k + 1
        k += 1
    raise ValueError(f"unbalanced {open_c} at {i}")


def _simple_end(src, i):
    """end of an empy 'simple expression' starting at an identifier: name(.name | [..] | (..))*"""
    n = len(src)
    k = i
    while k < n and (src[k].isalnum() or src[k] == "_"):
        k += 1
    while k < n:
        if src[k] == "." and k + 1 < n and (src[k + 1].isalpha() or src[k + 1] == "_"):
            k += 1
            while k < n and (src[k].isalnum() or src[k] == "_"):
                k += 1
        elif src[k] == "[":
            k = _match(src, k, "[", "]")
        elif src[k] == "(":
            k = _match(src, k, "(", ")")
        else:
            break
    return k


def _tokens(src):
    """[('text', s) | ('expr', code) | ('stmt', code) | ('ctl', code)]"""
    out, i, n, buf = [], 0, len(src), []
    while i < n:
        c = src[i]
        if c != "@":
            buf.append(c); i += 1
            continue
        nx = src[i + 1] if i + 1 < n else ""
        if nx == "@":
            buf.append("@"); i += 2
        elif nx in " \t\n":
            i += 2                                   # whitespace / line-continuation markup: consumed
        elif nx in "{[(":
            close = {"{": "}", "[": "]", "(": ")"}[nx]
            j = _match(src, i + 1, nx, close)
            if buf:
                out.append(("text", "".join(buf))); buf = []
            out.append(({"{": "stmt", "[": "ctl", "(": "expr"}[nx], src[i + 2:j - 1]))
            i = j
        elif nx.isalpha() or nx == "_":
            j = _simple_end(src, i + 1)
            if buf:
                out.append(("text", "".join(buf))); buf = []
            out.append(("expr", src[i + 1:j]))
            i = j
        else:
            raise ValueError(f"unsupported markup @{nx!r} at {i}")
    if buf:
        out.append(("text", "".join(buf)))
    return out


def _run(tokens, i, env, out, stop=None):
    while i < len(tokens):
        kind, val = tokens[i]
        if kind == "text":
            out.append(val)
        elif kind == "expr":
            out.append(str(eval(val, env)))
        elif kind == "stmt":
            exec(val, env)
        else:
            head = val.strip()
            if head.startswith("end"):
                assert stop is not None and head.split()[1] == stop, head
                return i
            assert head.startswith("for "), head
            target, iterable = head[4:].split(" in ", 1)
            # find the matching end
            depth, j = 0, i + 1
            while True:
                k, v = tokens[j]
                if k == "ctl":
                    h = v.strip()
                    if h.startswith("for "):
                        depth += 1
                    elif h.startswith("end"):
                        if depth == 0:
                            break
                        depth -= 1
                j += 1
            for item in eval(iterable, env):
                env["__item"] = item
                exec(f"{target.strip()} = __item", env)
                _run(tokens, i + 1, env, out, stop="for")
            i = j
        i += 1
    return i


def expand(template_text, local_vars):
    env = dict(local_vars)
    out = []
    _run(_tokens(template_text), 0, env, out)
    return "".join(out)


# ------------------------------------------------------------------------------------------- canonical form
def canonical(xml_text):
    """One line per element: depth, tag, attributes sorted by name with whitespace-normalised values."""
    root = ET.fromstring(xml_text)
    lines = []

    def walk(e, d):
        attrs = " ".join(f'{k}="{" ".join(v.split())}"' for k, v in sorted(e.attrib.items()))
        lines.append(f"{'  ' * d}{e.tag} {attrs}".rstrip())
        for c in e:
            walk(c, d + 1)
    walk(root, 0)
    return "\n".join(lines) + "\n"


def main():
    sys.path.insert(0, REF)
    from ft_grandprix.chunk import chunk
    from ft_grandprix.colors import resolve_color
    summary = {}
    rendered_hashes = {}
    cwd = os.getcwd()
    for track in ("track", "circle"):
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            chunk(os.path.join(REF, "template", f"{track}.png"), verbose=False, force=True, scale=2.0)   # custom.py:1155
            os.chdir(cwd)
            metadata = json.load(open(os.path.join(tmp, "rendered", "chunks", "metadata.json")))
            files = sorted(os.listdir(os.path.join(tmp, "rendered", "chunks")))
            rendered_hashes[track] = {
                "metadata": metadata if track == "circle" else {k: v for k, v in metadata.items() if k != "chunks"},
                "nfiles": len(files),
                "sha256_of_all_pngs": hashlib.sha256(b"".join(
                    f.encode() + open(os.path.join(tmp, "rendered", "chunks", f), "rb").read()
                    for f in files if f.endswith(".png"))).hexdigest(),
                "metadata_sha256": hashlib.sha256(open(os.path.join(tmp, "rendered", "chunks", "metadata.json"), "rb").read()).hexdigest(),
            }
        for template, cars_file, map_color in (("mushr", "cars", [1, 0, 0]), ("car", "cars", None), ("mushr", "all", [1, 0, 0])):
            if track == "circle" and (template, cars_file) != ("mushr", "cars"):
                continue
            cars = json.load(open(os.path.join(REF, "template", "cars", f"{cars_file}.json")))
            for index, car in enumerate(cars):                            # map.py:37-45
                for color in ("primary", "secondary"):
                    car[color] = resolve_color(car[color])
                car["x"] = 4.5 + 5.5 + 0.1 * (index % 3)
                car["y"] = -8.5 + 0.0 + 0.1 * (index % 3)
                car["z"] = 0.1
            text = open(os.path.join(REF, "template", f"{template}.em.xml")).read()
            xml = expand(text, {"map_color": map_color if map_color is not None else [1, 0, 0], "cars": cars,
                                "metadata": metadata, "rangefinders": 90, "scale": metadata["scale"]})   # map.py:58-64
            canon = canonical(xml)
            name = f"mjcf_{template}_{cars_file}_{track}.txt.gz"
            with gzip.GzipFile(os.path.join(HERE, name), "wb", mtime=0) as f:
                f.write(canon.encode())
            summary[name] = {"elements": canon.count("\n"), "sha256": hashlib.sha256(canon.encode()).hexdigest(),
                             "car_json": {"cars": cars, "rangefinders": 90}}
    json.dump({"mjcf": summary, "chunks": rendered_hashes}, open(os.path.join(HERE, "mjcf_golden.json"), "w"), indent=1)
    print(json.dumps({k: v["elements"] for k, v in summary.items()}))


if __name__ == "__main__":
    main()
