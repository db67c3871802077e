"""GeneratorQueue: the background-thread batch queue behind fit_generator(max_queue_size=, workers=) -- Keras'
GeneratorEnqueuer as the reference uses it (text_generation_model.py:470-472: max_queue_size=100, one worker)."""
import threading
import time

import pytest

from image_captioning_b200.text_model import GeneratorQueue


def _count(n, seen=None, delay=0.0):
    for i in range(n):
        if delay:
            time.sleep(delay)
        if seen is not None:
            seen.append((i, threading.current_thread().name))
        yield ([i, -i], i * i)


@pytest.mark.parametrize("workers", [0, 1, 4])
def test_batches_arrive_in_generator_order_and_stop_iteration_surfaces(workers):
    with GeneratorQueue(_count(7), max_queue_size=3, workers=workers) as q:
        assert [q.get() for _ in range(7)] == [([i, -i], i * i) for i in range(7)]
        for _ in range(2):                                  # and keeps surfacing
            with pytest.raises(StopIteration):
                q.get()


def test_worker_thread_runs_ahead_but_never_past_the_queue_bound():
    seen = []
    q = GeneratorQueue(_count(100, seen), max_queue_size=4, workers=1)
    time.sleep(0.3)
    # 4 queued + the one the worker holds while the queue is full
    assert 4 <= len(seen) <= 5, len(seen)
    assert all(name == "dcap-generator-queue" for _, name in seen)
    assert q.get() == ([0, 0], 0)
    time.sleep(0.2)
    assert len(seen) <= 6
    t0 = time.time()
    q.close()
    assert time.time() - t0 < 1.0
    n = len(seen)
    time.sleep(0.1)
    assert len(seen) == n                                   # the worker has stopped pulling


def test_generator_work_overlaps_the_consumer():
    """20 ms per batch in the generator, 20 ms per batch in the consumer: ~20 ms per step with the queue, ~40 ms without."""
    def run(workers):
        with GeneratorQueue(_count(10, delay=0.02), max_queue_size=4, workers=workers) as q:
            t0 = time.time()
            for _ in range(10):
                q.get()
                time.sleep(0.02)                            # "the training step" (a device wait releases the GIL like this)
            return time.time() - t0
    assert run(1) < 0.8 * run(0)


def test_generator_exception_reaches_the_consumer():
    def bad():
        yield 1
        raise ValueError("boom")
    with GeneratorQueue(bad(), workers=1) as q:
        assert q.get() == 1
        with pytest.raises(ValueError, match="boom"):
            q.get()
