"""The reference's interactive frame loop, headless: keyboard-driven camera, progressive accumulation that
restarts on every camera move, one dispatch per frame, `ms/frame` statistics.

  Camera            src/utils/Camera.h:28-134   (only .Position reaches the shader, main.cpp:174; yaw 180, pitch 0 at start)
  processInput      src/main.cpp:462-486        one direction per frame, the LAST pressed key in the order UP, DOWN, W, S, A, D wins
  updateScene       src/main.cpp:166-183        hasMoved -> currentSample = 0; write the 32-byte UBO; currentSample++
  drawFrame         src/main.cpp:323-395        -> ComputeModel.computeCommand(cmd, frame, W/32, H/32, 1)
  mainLoop          src/main.cpp:399-421        prints "%f ms/frame" once per second of frames
  frames in flight  src/main.cpp:68, :298-316, :325, :394   MAX_FRAMES_IN_FLIGHT = 2 slots, each behind a fence: FrameLoop(frames_in_flight=2)

Key presses come from a script (a string, one character per frame: w a s d u(p) j(down) . = no key) instead of GLFW.
"""
import math
import time

import numpy as np

from .scene import CAMERA_START, pack_ubo

FORWARD, BACKWARD, LEFT, RIGHT, UP, DOWN, NONE = "w", "s", "a", "d", "u", "j", "."


class Camera:
    """Camera.h:28-134 with the reference's defaults (YAW 180, PITCH 0, SPEED 2.5); fp32 arithmetic like glm's."""

    def __init__(self, position=CAMERA_START, up=(0.0, 1.0, 0.0), yaw=180.0, pitch=0.0):
        f = np.float32
        self.Position = np.array(position, f)
        self.WorldUp = np.array(up, f)
        self.Yaw, self.Pitch, self.MovementSpeed = f(yaw), f(pitch), f(2.5)
        self.updateCameraVectors()

    def updateCameraVectors(self):
        f = np.float32
        yaw, pitch = f(math.radians(self.Yaw)), f(math.radians(self.Pitch))
        front = np.array([np.cos(yaw) * np.cos(pitch), np.sin(pitch), np.sin(yaw) * np.cos(pitch)], f)
        self.Front = (front / f(np.sqrt(np.dot(front, front)))).astype(f)
        right = np.cross(self.Front, self.WorldUp).astype(f)
        self.Right = (right / f(np.sqrt(np.dot(right, right)))).astype(f)
        up = np.cross(self.Right, self.Front).astype(f)
        self.Up = (up / f(np.sqrt(np.dot(up, up)))).astype(f)

    def ProcessKeyboard(self, direction, deltaTime):
        velocity = np.float32(self.MovementSpeed * np.float32(deltaTime))
        step = {FORWARD: self.Front, BACKWARD: -self.Front, LEFT: -self.Right, RIGHT: self.Right, UP: self.Up, DOWN: -self.Up}.get(direction)
        if step is not None:
            self.Position = (self.Position + step * velocity).astype(np.float32)


class FrameLoop:
    """mainLoop/drawFrame of main.cpp around a ComputeModel.  `frame_time` is the deltaTime fed to the camera (the reference
    uses wall-clock frame time; a fixed value makes scripted runs reproducible)."""

    def __init__(self, model, scene, width, height, camera=None, frame_time=1.0 / 60.0, full_cover=True, frames_in_flight=0, on_present=None):
        """frames_in_flight = 0: every drawFrame is a synchronous-interface computeCommand (asynchronous on the context's one stream;
        nothing is read back until run() returns).  frames_in_flight = n >= 1: the reference's pipelined loop -- drawFrame waits for
        the fence of the slot it reuses (main.cpp:325), submits the frame on that slot's stream and moves on (:394); every frame's
        rgba8 image lands in the slot's page-locked buffer and `on_present(frame_index, array)` is called with it once its fence has
        signalled (the presentation of main.cpp:382-392), in frame order.  Frames are bit-identical either way."""
        self.model, self.scene = model, scene
        self.frames_in_flight = int(frames_in_flight)
        self.on_present = on_present
        self._slots = []            # per slot: [PinnedFrame, index of the frame in flight on it or None]
        self._width, self._height = width, height
        self.camera = camera or Camera()
        self.frame_time = frame_time
        self.gx = (width + 31) // 32 if full_cover else width // 32      # main.cpp:228 dispatches floor(W/32) x floor(H/32)
        self.gy = (height + 31) // 32 if full_cover else height // 32
        self.currentSample = 0
        self.frames = 0
        self.ms_per_frame = []          # one entry per wall-clock second, like the reference's printf (main.cpp:406-412)
        self._t_last, self._n_last = None, 0

    def drawFrame(self, key=NONE):
        hasMoved = key != NONE
        if hasMoved:                                            # processInput, main.cpp:481-485
            self.camera.ProcessKeyboard(key, self.frame_time)
            self.currentSample = 0                              # updateScene, main.cpp:169-173
        ubo = self.model.getMaterial().getUniformBufferBundles()[0].data.buffers[0]
        ubo.write(pack_ubo(tuple(float(x) for x in self.camera.Position), self.currentSample, self.scene, time=0.0))
        self.currentSample += 1                                 # main.cpp:182
        if self.frames_in_flight:
            self._draw_in_flight()
        else:
            self.model.computeCommand(None, 0, self.gx, self.gy, 1)
        self.frames += 1
        now = time.perf_counter()
        if self._t_last is None:
            self._t_last = now
        self._n_last += 1
        if now - self._t_last >= 1.0:
            self.ms_per_frame.append(1000.0 * (now - self._t_last) / self._n_last)
            self._t_last, self._n_last = now, 0

    # ---- the pipelined loop (frames_in_flight >= 1)
    def _begin(self):
        from .api import PinnedFrame
        self.model.getMaterial().framesBegin(self.frames_in_flight)
        self._slots = [[PinnedFrame(self._width, self._height), None] for _ in range(self.frames_in_flight)]
        self._next = 0

    def _retire(self, slot):
        buf, idx = self._slots[slot]
        if idx is not None:
            self.model.getMaterial().frameWait(slot)            # vkWaitForFences(inFlightFences[currentFrame]), main.cpp:325
            if self.on_present:
                self.on_present(idx, buf.array)
            self._slots[slot][1] = None

    def _draw_in_flight(self):
        if not self._slots:
            self._begin()
        slot = self._next
        self._retire(slot)
        got = self.model.frameCommand(None, 0, self.gx, self.gy, 1, out=self._slots[slot][0])
        assert got == slot
        self._slots[slot][1] = self.frames
        self._next = (slot + 1) % self.frames_in_flight         # main.cpp:394

    def finish(self):
        """vkDeviceWaitIdle (main.cpp:419): retires the frames still in flight, in order, and leaves the pipelined mode."""
        if self._slots:
            for k in range(self.frames_in_flight):
                self._retire((self._next + k) % self.frames_in_flight)
            self.model.getMaterial().framesEnd()
            for buf, _ in self._slots:
                buf.free()
            self._slots = []

    def run(self, script):
        """One frame per character of `script`; returns the final target image (H, W, 4) uint8."""
        for key in script:
            if key not in (FORWARD, BACKWARD, LEFT, RIGHT, UP, DOWN, NONE):
                raise ValueError("unknown key %r in camera script" % key)
            self.drawFrame(key)
        self.finish()
        return self.model.getMaterial().getStorageImages()[0].data.read()
