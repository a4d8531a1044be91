#!/usr/bin/env python
"""Convert the reference's ``maps_templates/*.json`` into this package's normalised ``catmap-1``
files under ``as_cops_and_thieves_b200/maps_data/``.

Run here (the container that has ``/root/reference``); the outputs are committed so that tests
and benchmarks on the GPU box never read ``/root/reference``.  The normalised form stores what the
reference ``Map`` *builds* from the file (``/root/reference/src/maps/map.py:35-117``): the closed
vertex ring of every block after the rect rule, and per agent its type, default position and
list of spawn regions.
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from as_cops_and_thieves_b200.maps import _parse_block, MAPS_DATA_DIR  # noqa: E402

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/maps_templates")


def convert(path: Path) -> dict:
    d = json.load(open(path))
    out = {
        "format": "catmap-1",
        "name": path.stem,
        "source": f"maps_templates/{path.name}",
        "window": list(d["window"].values()),
        "canvas": list(d["canvas"].values()),
        "blocks": [[c for v in _parse_block(b) for c in v] for b in d["objects"]["blocks"]],
    }
    agents = d.get("agents")
    if agents is None:
        out["agents"] = None
    else:
        out["agents"] = []
        for a in agents:
            regs = a.get("spawn_regions", a.get("spawn_region"))
            if isinstance(regs, dict):
                regs = [regs]
            out["agents"].append({
                "type": a["type"], "pos": [a["x"], a["y"]],
                "regions": [[r["x"], r["y"], r["w"], r["h"]] for r in (regs or [])],
            })
    return out


if __name__ == "__main__":
    MAPS_DATA_DIR.mkdir(exist_ok=True)
    for p in sorted(SRC.glob("*.json")):
        out = convert(p)
        dst = MAPS_DATA_DIR / f"{p.stem}.catmap.json"
        with open(dst, "w") as f:
            json.dump(out, f, separators=(",", ":"))
            f.write("\n")
        print(f"{p.name}: {len(out['blocks'])} blocks -> {dst}")
