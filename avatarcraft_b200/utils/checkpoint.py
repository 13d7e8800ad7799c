"""Checkpoint / resume and off-critical-path image writing.

Wire format: the reference saves `torch.save(net_style.state_dict(), <name>.pth.tar)` (stylize.py:255-260) and every
entry point loads exactly that.  `save_checkpoint` writes the same file -- so reference tools keep working -- and,
next to it, `<name>.resume.pt` with what the reference cannot restore: optimizer moments and step count, the global
step / epoch, and the CPU + CUDA RNG states.  `load_checkpoint` accepts either.

`AsyncImageWriter` moves PNG/GIF encoding off the render loop: the frame is copied device->pinned host on a side
stream, and a worker thread waits on the copy's event and encodes (the reference encodes synchronously per frame,
render_canonical.py:82-123, render_warp.py:107-122)."""
import os
import queue
import threading

import numpy as np
import torch


def save_checkpoint(path, net, optimizer=None, step=0, epoch=0, extra=None):
    """path: the reference-format weights file (`*.pth.tar`).  Returns the resume-file path (or None)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, path)
    if optimizer is None:
        return None
    resume = {"step": int(step), "epoch": int(epoch), "optimizer": optimizer.state_dict(),
              "rng_cpu": torch.get_rng_state(), "extra": extra or {}}
    if torch.cuda.is_available():
        resume["rng_cuda"] = torch.cuda.get_rng_state()
    rpath = resume_path(path)
    torch.save(resume, rpath)
    return rpath


def resume_path(path):
    return (path[:-len(".pth.tar")] if path.endswith(".pth.tar") else path) + ".resume.pt"


def load_checkpoint(path, net, optimizer=None, restore_rng=True):
    """Loads the weights (reference layout) and, when `<name>.resume.pt` exists and an optimizer is given, the
    optimizer state, counters and RNG streams.  Returns dict(step, epoch, extra, resumed)."""
    net.load_state_dict(torch.load(path, map_location="cpu"))
    info = {"step": 0, "epoch": 0, "extra": {}, "resumed": False}
    rpath = resume_path(path)
    if optimizer is not None and os.path.exists(rpath):
        r = torch.load(rpath, map_location="cpu", weights_only=False)
        optimizer.load_state_dict(r["optimizer"])
        if restore_rng:
            torch.set_rng_state(r["rng_cpu"])
            if "rng_cuda" in r and torch.cuda.is_available():
                torch.cuda.set_rng_state(r["rng_cuda"])
        info.update(step=r["step"], epoch=r["epoch"], extra=r.get("extra", {}), resumed=True)
    return info


def to_uint8(img):
    return (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


class AsyncImageWriter:
    """writer.submit(rgb_hw3_device_tensor, "out/frame_0001.png"); writer.close() joins (and writes the GIF)."""

    def __init__(self, gif_path=None, gif_duration_ms=100, slots=4):
        self._q = queue.Queue()
        self._gif_path, self._gif_ms, self._frames = gif_path, gif_duration_ms, []
        self._free = queue.Queue()
        self._slots = slots
        self._copy_stream = None
        self._err = None
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def submit(self, img, path):
        """img: [H,W,3] float tensor in [0,1] (CUDA or CPU).  Returns immediately after enqueueing the D2H copy."""
        if self._err:
            raise self._err
        if not img.is_cuda:
            self._q.put((img.detach().float().contiguous(), None, path, None))
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=img.device)
            for _ in range(self._slots):
                self._free.put(torch.empty(img.shape, dtype=torch.float32).pin_memory())
        host = self._free.get()                       # back-pressure: at most `slots` frames in flight
        if host.shape != img.shape:
            host = torch.empty(img.shape, dtype=torch.float32).pin_memory()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(img.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            img.record_stream(self._copy_stream)
            host.copy_(img.detach().float(), non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._q.put((host, done, path, host))

    def _work(self):
        from PIL import Image
        while True:
            item = self._q.get()
            if item is None:
                return
            host, done, path, slot = item
            try:
                if done is not None:
                    done.synchronize()
                frame = Image.fromarray(to_uint8(host.numpy()))
                os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
                frame.save(path)
                if self._gif_path:
                    self._frames.append(frame)
            except Exception as e:                    # surfaced on the next submit / close
                self._err = e
            finally:
                if slot is not None:
                    self._free.put(slot)

    def close(self):
        self._q.put(None)
        self._t.join()
        if self._err:
            raise self._err
        if self._gif_path and self._frames:
            self._frames[0].save(self._gif_path, save_all=True, append_images=self._frames[1:], duration=self._gif_ms, loop=0)
