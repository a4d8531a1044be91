#!/usr/bin/env python
"""Author the analytic known-answer vectors (tests/golden/analytic_vectors.json).

The reference ships no tests or golden vectors and pymunk cannot be installed here (SURVEY.md §4,
§8c), so these are derived from first principles — closed-form geometry and the published
Chipmunk2D contact equations — NOT by running the oracle or the CUDA path.  Every expectation
below is a formula evaluated in this file.

Map: three axis-aligned blocks (reference schema, tests/golden/analytic_map.json)
  A: x,y in [100,200]            the wall the rays and the contact tests use
  B: x in [300,304], y in [95,155]   thin wall for the line-of-sight test
  C: x,y in [600,610]            far away (keeps the static index a real tree)
Agents: cop_0, cop_1, thief_0.  Radii: agent 5, wall 1, ray 1; ray range 400; dt 1/60.
"""
import json
import math
from pathlib import Path

HERE = Path(__file__).resolve().parent
DT = 1.0 / 60.0
FAR = [[900.0, 700.0], [950.0, 700.0], [1000.0, 700.0]]  # parking spots, > 400 + from everything used


def ray_dir(i, n=90):
    a = i * (2.0 * math.pi / n)
    return math.cos(a), math.sin(a)


vectors = []

# (1) ray perpendicular to an axis-aligned wall: hit point on the rounded surface, distance D - 1
# origin (50,150), ray 0 (+x).  Raw face x=100 at D=50.  Fat ray (r=1) centre stops at 100-2; the
# reported point is centre - n*r_ray with n=(-1,0) -> x = 99.
vectors.append(dict(name="ray_perpendicular_wall", kind="ray",
                    pos=[[50.0, 150.0], FAR[1], FAR[2]], agent=0, ray=0,
                    expect=dict(type=0, point=[99.0, 150.0], distance=49.0)))
# same wall seen from the right: origin (260,150), ray 45 (180 deg, -x): face x=200 -> point x=201
vectors.append(dict(name="ray_perpendicular_wall_negx", kind="ray",
                    pos=[[260.0, 150.0], FAR[1], FAR[2]], agent=0, ray=45,
                    expect=dict(type=0, point=[201.0, 150.0], distance=59.0)))

# (2) ray passing a hull corner inside the shape's bb: bevel circle centre = vertex, radius 1+1.
# origin (50, 99.5), ray +x.  Corner (100,100): |P - corner| = 2 with P=(x,99.5) ->
# x = 100 - sqrt(4 - 0.25); n = (P - corner)/2; point = P - n*1.
px = 100.0 - math.sqrt(4.0 - 0.25)
nx, ny = (px - 100.0) / 2.0, (99.5 - 100.0) / 2.0
vectors.append(dict(name="ray_corner_bevel", kind="ray",
                    pos=[[50.0, 99.5], FAR[1], FAR[2]], agent=0, ray=0,
                    expect=dict(type=0, point=[px - nx, 99.5 - ny],
                                distance=math.hypot(px - nx - 50.0, 99.5 - ny - 99.5))))
# (2b) same corner but the THIN ray (y=98.5) stays outside the shape bb [99,201]^2: the BB-tree never
# visits the hull (the spatial index ignores the query radius) although the fat ray would touch it.
# The ray therefore carries on to wall B (face x=300): point x = 299, distance 249.
vectors.append(dict(name="ray_corner_outside_bb_misses", kind="ray",
                    pos=[[50.0, 98.5], FAR[1], FAR[2]], agent=0, ray=0,
                    expect=dict(type=0, point=[299.0, 98.5], distance=249.0)))
# ... and looking the other way from the far side nothing at all is hit through that corner sliver
vectors.append(dict(name="ray_corner_outside_bb_misses_empty", kind="ray",
                    pos=[[250.0, 201.5], FAR[1], FAR[2]], agent=0, ray=45,
                    expect=dict(type=4, distance=400.0)))

# (3) ray to another agent, head-on: first contact of the fat ray with a circle of radius 5+1.
# cop_0 (50,50) looks +x at cop_1 (80,50): centre stops at 80-6, point = 75 -> distance L-5 = 25.
vectors.append(dict(name="ray_hits_cop", kind="ray",
                    pos=[[50.0, 50.0], [80.0, 50.0], FAR[2]], agent=0, ray=0,
                    expect=dict(type=1, point=[75.0, 50.0], distance=25.0)))
# thief seen by a cop -> type THIEF (2); cop seen by the thief -> COP (1)
vectors.append(dict(name="ray_hits_thief", kind="ray",
                    pos=[[50.0, 50.0], FAR[1], [80.0, 50.0]], agent=0, ray=0,
                    expect=dict(type=2, point=[75.0, 50.0], distance=25.0)))
vectors.append(dict(name="thief_sees_cop", kind="ray",
                    pos=[[50.0, 50.0], FAR[1], [80.0, 50.0]], agent=2, ray=45,
                    expect=dict(type=1, point=[55.0, 50.0], distance=25.0)))
# off-axis circle hit: origin (50,50), target centre (80,53), ray +x: P.x = 80 - sqrt(36-9)
s = 80.0 - math.sqrt(36.0 - 9.0)
cnx, cny = (s - 80.0) / 6.0, (50.0 - 53.0) / 6.0
vectors.append(dict(name="ray_hits_cop_offaxis", kind="ray",
                    pos=[[50.0, 50.0], [80.0, 53.0], FAR[2]], agent=0, ray=0,
                    expect=dict(type=1, point=[s - cnx, 50.0 - cny],
                                distance=math.hypot(s - cnx - 50.0, -cny))))

# (4) nothing within range -> (400, EMPTY)
vectors.append(dict(name="ray_no_hit", kind="ray",
                    pos=[[50.0, 50.0], FAR[1], FAR[2]], agent=0, ray=45,
                    expect=dict(type=4, distance=400.0)))
# wall farther than the range: origin (50,400) ray +x sees C? no: C is at y 600; nothing -> EMPTY
vectors.append(dict(name="ray_range_limit", kind="ray",
                    pos=[[1010.5, 605.0], FAR[1], FAR[2]], agent=0, ray=45,
                    # C's face x=610 is 400.5 away: fat centre would stop at 612 -> s = 398.5 < 400 -> HIT
                    expect=dict(type=0, point=[611.0, 605.0], distance=399.5)))

# (5) free flight: v' = v + dv(action), clamp at 125; p' = p + v' * dt
def free(pos, vel, act):
    dv = {0: (-10.0, 0.0), 1: (0.0, 10.0), 2: (10.0, 0.0), 3: (0.0, -10.0)}[act]
    vx, vy = vel[0] + dv[0], vel[1] + dv[1]
    sp = math.hypot(vx, vy)
    if sp > 125.0:
        vx, vy = vx / sp * 125.0, vy / sp * 125.0
    return [pos[0] + vx * DT, pos[1] + vy * DT], [vx, vy]


p0, v0 = free([400.0, 400.0], [3.0, -4.0], 2)
p1, v1 = free([450.0, 450.0], [0.0, 120.0], 1)      # 130 -> clamped to 125
p2, v2 = free([500.0, 400.0], [100.0, 100.0], 0)    # |(90,100)| = 134.5 -> clamped, direction kept
vectors.append(dict(name="free_flight_and_clamp", kind="step",
                    pos=[[400.0, 400.0], [450.0, 450.0], [500.0, 400.0]],
                    vel=[[3.0, -4.0], [0.0, 120.0], [100.0, 100.0]], actions=[2, 1, 0],
                    expect=dict(pos=[p0, p1, p2], vel=[v0, v1, v2], vbias=[[0, 0]] * 3,
                                terminated=0, truncated=0, winner=-1)))

# (6) head-on wall contact (e=0, mu=0): normal velocity -> 0, tangential unchanged;
# bias velocity = 0.1 * (pen - slop) / dt with biasCoef = 1 - 0.9^(60 dt) = 0.1.
# cop_0 at (94.5,150), v=(30,7), action right -> (40,7).  Integrate: x = 94.5 + 40/60.
x1 = 94.5 + 40.0 * DT
pen = 6.0 - (100.0 - x1)
vb = 0.1 * (pen - 0.1) / DT
vectors.append(dict(name="wall_contact_head_on", kind="step",
                    pos=[[94.5, 150.0], FAR[1], FAR[2]], vel=[[30.0, 7.0], [0.0, 0.0], [0.0, 0.0]],
                    actions=[2, 1, 1], check_agents=[0],
                    expect=dict(pos=[[x1, 150.0 + 7.0 * DT]], vel=[[0.0, 7.0]], vbias=[[-vb, 0.0]],
                                terminated=0, truncated=0, winner=-1)))
# touching but not penetrating beyond the slop: contact exists (d <= 6), bias 0, approach velocity removed
x2 = 93.0 + 70.0 * DT     # = 94.1667 -> d = 5.8333 <= 6, pen 0.1667 - slop 0.1 -> small bias
pen2 = 6.0 - (100.0 - x2)
vectors.append(dict(name="wall_contact_shallow", kind="step",
                    pos=[[93.0, 150.0], FAR[1], FAR[2]], vel=[[60.0, 0.0], [0.0, 0.0], [0.0, 0.0]],
                    actions=[2, 1, 1], check_agents=[0],
                    expect=dict(pos=[[x2, 150.0]], vel=[[0.0, 0.0]], vbias=[[-0.1 * max(0.0, pen2 - 0.1) / DT, 0.0]],
                                terminated=0, truncated=0, winner=-1)))
# no contact when the centre stays farther than 6 from the hull: velocity kept
vectors.append(dict(name="wall_no_contact", kind="step",
                    pos=[[90.0, 150.0], FAR[1], FAR[2]], vel=[[20.0, 0.0], [0.0, 0.0], [0.0, 0.0]],
                    actions=[2, 1, 1], check_agents=[0],
                    expect=dict(pos=[[90.0 + 30.0 * DT, 150.0]], vel=[[30.0, 0.0]], vbias=[[0.0, 0.0]],
                                terminated=0, truncated=0, winner=-1)))

# (7) two equal circles head-on: relative normal velocity -> 0, momentum conserved.
# cop_0 (400,300) v (20,0) +right -> 30 ; cop_1 (409,300) v (-10,0) +left -> -20.
xa, xb = 400.0 + 30.0 * DT, 409.0 - 20.0 * DT
dist = xb - xa
j = 0.5 * 50.0            # nMass 1/2 * closing speed 50
bias = 0.1 * max(0.0, (10.0 - dist) - 0.1) / DT
vectors.append(dict(name="two_circles_head_on", kind="step",
                    pos=[[400.0, 300.0], [409.0, 300.0], FAR[2]], vel=[[20.0, 0.0], [-10.0, 0.0], [0.0, 0.0]],
                    actions=[2, 0, 1], check_agents=[0, 1],
                    expect=dict(pos=[[xa, 300.0], [xb, 300.0]], vel=[[30.0 - j, 0.0], [-20.0 + j, 0.0]],
                                vbias=[[-bias / 2.0, 0.0], [bias / 2.0, 0.0]],
                                terminated=0, truncated=0, winner=-1)))

# (8) capture / line of sight / timeout
# thief 15 away from cop_0 with clear line of sight -> captured on this call (pre-step state)
vectors.append(dict(name="capture_clear_los", kind="step",
                    pos=[[400.0, 500.0], FAR[1], [415.0, 500.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=1, truncated=0, winner=0, reward=[1.0, 1.0, -1.0])))
# exactly at the radius: dist < 20 is strict -> no capture
vectors.append(dict(name="capture_radius_strict", kind="step",
                    pos=[[400.0, 500.0], FAR[1], [420.0, 500.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=0, truncated=0, winner=-1)))
# thin wall B between them (cop at x=293, thief at x=311, 18 apart): walls block -> no capture
vectors.append(dict(name="capture_blocked_by_wall", kind="step",
                    pos=[[293.0, 120.0], FAR[1], [311.0, 120.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=0, truncated=0, winner=-1)))
# another COP standing between them does not block (both agent categories are masked out)
vectors.append(dict(name="capture_not_blocked_by_agent", kind="step",
                    pos=[[400.0, 500.0], [409.0, 500.0], [418.0, 500.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=1, truncated=0, winner=0, reward=[1.0, 1.0, -1.0])))
# timeout: the call on which step_count reaches max_step_count -> thief wins, truncations set
vectors.append(dict(name="timeout_thief_wins", kind="step", step_count=399, max_step_count=400,
                    pos=[FAR[0], FAR[1], [100.0, 700.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=1, truncated=1, winner=1, reward=[-1.0, -1.0, 1.0])))
# capture wins over timeout
vectors.append(dict(name="capture_beats_timeout", kind="step", step_count=399, max_step_count=400,
                    pos=[[400.0, 500.0], FAR[1], [415.0, 500.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=1, truncated=0, winner=0, reward=[1.0, 1.0, -1.0])))

# rewards from the f16 distances (cop.py:66-72, thief.py:60-66), mid-episode
# cop_0 (50,50) sees thief at distance 25 (vector 3): -0.02 + 1.5*exp(-25/50); thief sees cop at 25:
# tanh((25-100)/50)/10; cop_1 far away sees nothing: -0.04
vectors.append(dict(name="rewards_mid_episode", kind="step",
                    pos=[[50.0, 50.0], FAR[1], [80.0, 50.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, capture_radius_override=1.0,
                    expect=dict(terminated=0, truncated=0, winner=-1,
                                reward=[-0.02 + 1.5 * math.exp(-25.0 / 50.0), -0.04,
                                        math.tanh((25.0 - 100.0) / 50.0) / 10.0])))
# thief sees no cop -> 0.15
vectors.append(dict(name="rewards_nothing_seen", kind="step",
                    pos=[FAR[0], FAR[1], [100.0, 700.0]], vel=[[0, 0]] * 3, actions=[1, 1, 1],
                    auto_reset=0, expect=dict(terminated=0, truncated=0, winner=-1, reward=[-0.04, -0.04, 0.15])))

analytic_map = {
    "window": {"w_px": 1280, "h_px": 800}, "canvas": {"w": 1280, "h": 800},
    "objects": {"blocks": [
        {"type": "rect", "x": 100, "y": 100, "w": 100, "h": 100},
        {"type": "rect", "x": 304, "y": 155, "w": -4, "h": -60},   # negative extents (map.py:44-52)
        {"type": "poly", "vs": [{"x": 600, "y": 600}, {"x": 610, "y": 600}, {"x": 610, "y": 610},
                                {"x": 605, "y": 610}, {"x": 600, "y": 610}, {"x": 600, "y": 600}]},
    ]},
    "agents": [
        {"type": "cop", "x": 900, "y": 700, "spawn_region": {"x": 880, "y": 680, "w": 40, "h": 40}},
        {"type": "thief", "x": 1000, "y": 700,
         "spawn_regions": [{"x": 980, "y": 680, "w": 40, "h": 40}, {"x": 700, "y": 100, "w": 50, "h": 50}]},
        {"type": "cop", "x": 950, "y": 700},
    ],
}

if __name__ == "__main__":
    with open(HERE / "analytic_map.json", "w") as f:
        json.dump(analytic_map, f, indent=1)
    with open(HERE / "analytic_vectors.json", "w") as f:
        json.dump({"dt": DT, "vectors": vectors}, f, indent=1)
    print(f"wrote {len(vectors)} vectors")
