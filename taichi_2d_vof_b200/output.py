"""The reference's on-disk output (`-s`: `output/%06d-f.png`, 2dvof.py:563-571; 3-D: `output/step-%05d.vtr`,
3dvof.py:624-627) without stalling the time loop.

The reference blocks on ``F.to_numpy()`` and on the file write every ``nstep`` steps.  Here the field is snapshotted
device-to-device on the compute stream (``vof*_field_get_async``), copied to a pinned host buffer on a side stream, and
written by a worker thread; the solver keeps stepping meanwhile.  Two pinned buffers alternate, so at most one copy and
one file write are in flight.
"""
from __future__ import annotations

import queue
import threading

from . import _lib


class FieldDumper:
    def __init__(self, shape, write):
        """``write(array, tag)`` is called on the worker thread once ``array`` holds the field."""
        self._bufs = [_lib.pinned_empty(shape), _lib.pinned_empty(shape)]
        self._free = [threading.Event(), threading.Event()]
        for e in self._free:
            e.set()
        self._k = 0
        self._copied = threading.Event()
        self._copied.set()
        self._write = write
        self._q = queue.Queue()
        self._err = None
        self._t = threading.Thread(target=self._worker, daemon=True)
        self._t.start()
        self.dumps = 0

    def _worker(self):
        while True:
            job = self._q.get()
            if job is None:
                return
            field, k, tag, copied = job
            try:
                field.wait()                      # the side-stream copy of this snapshot
                copied.set()
                self._write(self._bufs[k], tag)
            except Exception as e:                # surfaced by the next dump() / close()
                self._err = e
            finally:
                copied.set()
                self._free[k].set()
                self._q.task_done()

    def dump(self, field, tag):
        """Start a non-stalling read of ``field`` (a solver Field) and queue the file write."""
        if self._err:
            raise self._err
        k = self._k
        self._k ^= 1
        self._free[k].wait()                      # its previous file is on disk
        self._free[k].clear()
        self._copied.wait()                       # one read in flight per context: the previous copy has landed
        self._copied = threading.Event()
        field.to_numpy_async(self._bufs[k])
        self._q.put((field, k, tag, self._copied))
        self.dumps += 1

    def dump_array(self, arr, tag):
        """Queue the write of an array that is already on the host (multi-GPU gather)."""
        if self._err:
            raise self._err

        class _Ready:
            def wait(self_inner):
                pass
        k = self._k
        self._k ^= 1
        self._free[k].wait()
        self._free[k].clear()
        self._bufs[k][...] = arr
        self._q.put((_Ready(), k, tag, threading.Event()))
        self.dumps += 1

    def close(self):
        self._q.join()
        self._q.put(None)
        self._t.join()
        for b in self._bufs:
            _lib.pinned_free(b)
        self._bufs = []
        if self._err:
            raise self._err
