"""CPU model of the arrival / merge hand-shake of the fused decode attention (decode_attn_fused.cu): every item of a (row, kv head) pair
arrives on one counter; the SP private items each wait for ALL n_slots arrivals, merge their share, and the last of them to get past the
wait resets the counter and the "pollers done" word for the next launch.  Randomly interleaved: nobody may miss the full count, both words
must be zero when the launch ends, and a second launch on the same words must behave identically."""
import random

import pytest


def launch(state, n_shared, n_private, rng):
    n_slots = n_shared + n_private
    arrived = {"n": 0}
    agents = []

    def shared_item():
        yield                                   # tile work
        state["c"] += 1                          # red.release.gpu.add
        arrived["n"] += 1

    def private_item():
        yield
        state["c"] += 1
        arrived["n"] += 1
        while state["c"] < n_slots:              # ld.acquire poll; the counter only grows until the last poller resets it
            yield
        assert arrived["n"] == n_slots, "a poller passed before every partial was published"
        old = state["dn"]; state["dn"] += 1      # atomicAdd
        if old == n_private - 1:
            state["c"] = 0; state["dn"] = 0
        yield                                    # merge of this item's share

    agents = [shared_item() for _ in range(n_shared)] + [private_item() for _ in range(n_private)]
    live = list(range(len(agents)))
    steps = 0
    while live:
        i = rng.choice(live)
        try:
            next(agents[i])
        except StopIteration:
            live.remove(i)
        steps += 1
        assert steps < 100000, "merge hand-shake does not terminate"
    assert state == {"c": 0, "dn": 0}, state


@pytest.mark.parametrize("n_shared,n_private", [(14, 3), (8, 2), (0, 8), (28, 1), (1, 1)])
def test_decode_merge_handshake(n_shared, n_private):
    for seed in range(50):
        rng = random.Random(seed)
        state = {"c": 0, "dn": 0}
        for _ in range(3):                       # consecutive launches (kernel boundaries) reuse the same two words
            launch(state, n_shared, n_private, rng)
