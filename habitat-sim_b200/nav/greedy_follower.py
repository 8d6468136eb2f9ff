"""GreedyGeodesicFollower on top of the batched PathFinder.

Restates esp::nav::GreedyGeodesicFollowerImpl (src/esp/nav/GreedyFollower.cpp:38-240) and the
Python wrapper src_python/habitat_sim/nav/greedy_geodesic_follower.py:18-200.  The reference
evaluates each primitive `[LEFT]*n + [FORWARD]` / `[RIGHT]*n + [FORWARD]` by bouncing
C++ -> Python -> C++ through a MoveFn (GreedyFollower.h:54) and running try_step, find_path
and distance_to_closest_obstacle one at a time (GreedyFollower.cpp:46-81).  Here all
candidates of one decision are evaluated by ONE batched try_steps / find_paths /
distances_to_closest_obstacle call, and the reference's sequential selection (first strictly
better reward, early exit above 0.99, GreedyFollower.cpp:99-137) is replayed over the results.

The forward / turn kinematics are those of habitat_sim/agent/controls/default_controls.py
(move along local -Z, rotate about +Y) followed by ObjectControls.action's step filter and
collision test (object_controls.py:50-90).
"""
from __future__ import annotations

import enum
import math

import numpy as np


class GreedyFollowerCodes(enum.IntEnum):
    """GreedyFollower.h:39-45"""
    ERROR = -2
    STOP = -1
    FORWARD = 0
    LEFT = 1
    RIGHT = 2


_EPS = 1e-5  # object_controls.py: EPS
_CLOSE_TO_OBS = 0.2  # closeToObsThreshold_, GreedyFollower.h:129
_COLLISION_COST = 0.25  # collisionCost_, GreedyFollower.h:130


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], np.float64)


def _quat_rotate(q, v):
    x, y, z, w = q
    u = np.array([x, y, z])
    return v + 2.0 * np.cross(u, np.cross(u, v) + w * v)


def _yaw_quat(theta):
    return np.array([0.0, math.sin(theta / 2.0), 0.0, math.cos(theta / 2.0)], np.float64)


class GreedyGeodesicFollowerImpl:
    """SPB.cpp:287-312 constructor order: (pathfinder, move_forward, turn_left, turn_right,
    goal_dist, forward_amount, turn_amount, fix_thrashing, thrashing_threshold).  The three
    MoveFn arguments are accepted for signature compatibility; when they are None the
    default kinematics are used, which is what enables batched evaluation."""

    def __init__(self, pathfinder, move_forward=None, turn_left=None, turn_right=None,
                 goal_dist: float = 0.1875, forward_amount: float = 0.25,
                 turn_amount: float = math.radians(10.0), fix_thrashing: bool = True,
                 thrashing_threshold: int = 16):
        self.pathfinder = pathfinder
        self.goal_dist = float(goal_dist)
        self.forward_amount = float(forward_amount)
        self.turn_amount = float(turn_amount)
        self.fix_thrashing = bool(fix_thrashing)
        self.thrashing_threshold = int(thrashing_threshold)
        self._actions: list = []
        self._thrashing_actions: list = []

    def reset(self):
        self._actions.clear()
        self._thrashing_actions.clear()

    # -- kinematics ---------------------------------------------------------------------
    def _forward_target(self, rot, pos):
        return pos + _quat_rotate(rot, np.array([0.0, 0.0, -self.forward_amount]))

    def _turn(self, rot, sign):
        q = _quat_mul(rot, _yaw_quat(sign * self.turn_amount))
        return q / np.linalg.norm(q)

    # -- GreedyFollower.cpp:83-140 ---------------------------------------------------------
    def _next_best_prim_along(self, rot, pos, end, geo):
        if geo == math.inf:
            return [GreedyFollowerCodes.ERROR]
        if geo < self.goal_dist:
            return [GreedyFollowerCodes.STOP]
        # candidate headings in the reference's loop order: angle = 0, t, 2t, ... < pi (f32 accumulate)
        n_steps = 0
        angle = np.float32(0.0)
        while angle < np.float32(math.pi):
            n_steps += 1
            angle = np.float32(angle + np.float32(self.turn_amount))
        lrot, rrot = rot.copy(), rot.copy()
        targets, meta = [], []
        for k in range(n_steps):
            targets.append(self._forward_target(lrot, pos))
            meta.append((GreedyFollowerCodes.LEFT, k))
            targets.append(self._forward_target(rrot, pos))
            meta.append((GreedyFollowerCodes.RIGHT, k))
            lrot = self._turn(lrot, +1.0)
            rrot = self._turn(rrot, -1.0)
        targets = np.asarray(targets, np.float32)
        starts = np.repeat(np.asarray(pos, np.float32)[None], len(targets), 0)
        # one batched evaluation of every primitive (GreedyFollower.cpp:46-60)
        filt = self.pathfinder.try_steps(starts, targets, True)
        before = ((targets - starts).astype(np.float32) ** 2).sum(1)
        after = ((filt - starts).astype(np.float32) ** 2).sum(1)
        collided = (after + _EPS) < before
        geo_after = self.pathfinder.geodesic_distances(filt, np.repeat(np.asarray(end, np.float32)[None], len(filt), 0))
        obs_after = self.pathfinder.distances_to_closest_obstacle(filt, 1.1 * _CLOSE_TO_OBS)
        best_reward = -_COLLISION_COST
        best = []
        for i, (side, k) in enumerate(meta):
            reward = (np.float32(geo) - geo_after[i]) / np.float32(self.forward_amount) + (
                -0.0125 * k - (_COLLISION_COST if collided[i] else 0.0)
                - (0.05 if obs_after[i] < _CLOSE_TO_OBS else 0.0))
            if reward > best_reward:
                best_reward = reward
                best = [side] * k + [GreedyFollowerCodes.FORWARD]
            if side == GreedyFollowerCodes.RIGHT and best_reward > 0.99:
                break
        return best

    def _is_thrashing(self):
        if len(self._actions) < self.thrashing_threshold:
            return False
        last = self._actions[-1]
        thrashing = last in (GreedyFollowerCodes.LEFT, GreedyFollowerCodes.RIGHT)
        i = 2
        while i < self.thrashing_threshold + 1 and thrashing:
            a = self._actions[-i]
            thrashing = ((a == GreedyFollowerCodes.RIGHT and last == GreedyFollowerCodes.LEFT)
                         or (a == GreedyFollowerCodes.LEFT and last == GreedyFollowerCodes.RIGHT))
            last = a
            i += 1
        return thrashing

    def _geo(self, pos, end):
        return float(self.pathfinder.geodesic_distances(np.asarray(pos, np.float32)[None],
                                                        np.asarray(end, np.float32)[None])[0])

    # -- GreedyFollower.cpp:160-188 ----------------------------------------------------------
    def next_action_along(self, current_rot, current_pos, end):
        rot = np.asarray(current_rot, np.float64)
        pos = np.asarray(current_pos, np.float64)
        geo = self._geo(pos, end)
        if self.fix_thrashing and self._thrashing_actions:
            nxt = self._thrashing_actions.pop()
        else:
            acts = self._next_best_prim_along(rot, pos, end, geo)
            if not acts:
                nxt = GreedyFollowerCodes.ERROR
            elif self.fix_thrashing and self._is_thrashing():
                self._thrashing_actions = list(reversed(acts))
                nxt = self._thrashing_actions.pop()
            else:
                nxt = acts[0]
        self._actions.append(nxt)
        return nxt

    # -- GreedyFollower.cpp:190-240 ----------------------------------------------------------
    def find_path(self, current_rot, current_pos, end, max_actions: int = 5000):
        rot = np.asarray(current_rot, np.float64).copy()
        pos = np.asarray(current_pos, np.float64).copy()
        while True:
            geo = self._geo(pos, end)
            prim = self._next_best_prim_along(rot, pos, end, geo)
            if not prim:
                self._actions.append(GreedyFollowerCodes.ERROR)
            else:
                for a in prim:
                    if a == GreedyFollowerCodes.FORWARD:
                        tgt = self._forward_target(rot, pos)
                        pos = self.pathfinder.try_step(pos, tgt).astype(np.float64)
                    elif a == GreedyFollowerCodes.RIGHT:
                        rot = self._turn(rot, -1.0)
                    elif a == GreedyFollowerCodes.LEFT:
                        rot = self._turn(rot, +1.0)
                    self._actions.append(a)
            if (self._actions[-1] in (GreedyFollowerCodes.STOP, GreedyFollowerCodes.ERROR)
                    or len(self._actions) >= max_actions):
                break
        if self._actions[-1] == GreedyFollowerCodes.ERROR or len(self._actions) >= max_actions:
            return []
        return list(self._actions)


class GreedyGeodesicFollower:
    """habitat_sim.nav.GreedyGeodesicFollower twin that owns its (default) kinematics instead
    of an Agent: state is (rotation quaternion xyzw, position)."""

    def __init__(self, pathfinder, goal_radius=None, *, forward_amount: float = 0.25,
                 turn_degrees: float = 10.0, stop_key=None, forward_key="move_forward",
                 left_key="turn_left", right_key="turn_right", fix_thrashing: bool = True,
                 thrashing_threshold: int = 16):
        self.pathfinder = pathfinder
        self.goal_radius = 0.75 * forward_amount if goal_radius is None else goal_radius
        self.action_mapping = {GreedyFollowerCodes.STOP: stop_key,
                               GreedyFollowerCodes.FORWARD: forward_key,
                               GreedyFollowerCodes.LEFT: left_key,
                               GreedyFollowerCodes.RIGHT: right_key}
        self.impl = GreedyGeodesicFollowerImpl(pathfinder, None, None, None, self.goal_radius,
                                               forward_amount, math.radians(turn_degrees),
                                               fix_thrashing, thrashing_threshold)
        self.last_goal = None

    def reset(self):
        self.impl.reset()
        self.last_goal = None

    def next_action_along(self, rotation, position, goal_pos):
        if self.last_goal is None or not np.allclose(goal_pos, self.last_goal):
            self.reset()
            self.last_goal = np.asarray(goal_pos)
        act = self.impl.next_action_along(rotation, position, goal_pos)
        if act == GreedyFollowerCodes.ERROR:
            raise RuntimeError("GreedyFollowerError")
        return self.action_mapping[act]

    def find_path(self, rotation, position, goal_pos):
        self.reset()
        path = self.impl.find_path(rotation, position, goal_pos)
        if not path:
            raise RuntimeError("GreedyFollowerError")
        return [self.action_mapping[a] for a in path]
