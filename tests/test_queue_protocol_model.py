"""Model check of the finish-order queue protocol of the pipelined step launches (CPU only, no CUDA).

The step kernel's launch-to-launch hand-over (`gym_quadruped_b200/csrc/qs_kernel.cuh`: slot claim at the top of `env_kernel`, publish
block at its end; host side `step_impl` in `qstep.cu`) is restated here as a small discrete-event model and run under random
schedules: a ring of `D` queues, launches of `ceil(n / W)` CTAs of `W` warps, `S` SMs that hold one CTA each, programmatic
dependent launch (a launch may start once every CTA of the previous one has started), convoy start (CTA-wide barrier after the
slot claim), heavy-tailed per-env durations.  Checked on every run:

  * no deadlock: as long as work is left, some resident warp can make progress or a CTA can be placed;
  * every env is stepped by the launches 0, 1, 2, ... in this order, exactly once each;
  * every publish position lies inside the queue; publish counters and generation tags are modelled with small moduli so that
    both wrap many times.

Two mutations show that the model can see the two defects fixed late in round 2: without the generation
tag a younger launch takes an env published for an older one; without the publish-counter wait a position leaves the queue.
"""
import random

import pytest


class Violation(Exception):
    pass


def simulate(n, W, D, S, L, seed, check_generation=True, wait_for_counter=True, MOD=1 << 12, G=16, max_events=2_000_000):
    rnd = random.Random(seed)
    nct = (n + W - 1) // W
    # shared state: slot = ('F', gen, env) or ('E', gen)
    q = [[('F', 0, i) for i in range(n)]] + [[('E', (0 - 1) % G) for _ in range(n)] for _ in range(D - 1)]
    tails = [0] * D
    next_launch = [0] * n           # the launch that must step env e next
    started = [0] * L               # CTAs of launch s that have become resident so far
    resident = []                   # CTA records
    free_sm = S
    place = [0, 0]                  # next (launch, cta) to place
    done_envsteps = 0

    def signed(x):                  # difference of two MOD-counters, interpreted in (-MOD/2, MOD/2]
        x %= MOD
        return x - MOD if x >= MOD // 2 else x

    def new_cta(s, c):
        warps = []
        for w in range(W):
            slot = c * W + w
            warps.append(dict(state=0 if slot < n else 1, slot=slot, env=None, left=0, pos=None, active=slot < n))
        return dict(s=s, c=c, warps=warps, inp=s % D, out=(s + 1) % D, gen_in=(s // D) % G, gen_out=((s + 1) // D) % G, base=((s // D) * n) % MOD)

    def try_place():
        nonlocal free_sm
        progressed = False
        while free_sm > 0 and place[0] < L:
            s, c = place
            if s > 0 and started[s - 1] < nct:      # programmatic dependent launch: every CTA of the previous launch has started
                break
            resident.append(new_cta(s, c))
            started[s] += 1
            free_sm -= 1
            place[1] += 1
            if place[1] == nct:
                place[0] += 1; place[1] = 0
            progressed = True
        return progressed

    def step_warp(cta, wp):
        """One attempt of a warp; returns True if its state changed."""
        nonlocal done_envsteps
        st = wp['state']
        if st == 0:                                 # slot claim (qs_kernel.cuh, top of env_kernel)
            v = q[cta['inp']][wp['slot']]
            if v[0] == 'F' and (v[1] == cta['gen_in'] or not check_generation):
                env = v[2]
                if next_launch[env] != cta['s']:
                    raise Violation(f"env {env} taken by launch {cta['s']} but launch {next_launch[env]} is due")
                q[cta['inp']][wp['slot']] = ('E', cta['gen_in'])
                wp['env'] = env
                wp['state'] = 1
                return True
            return False
        if st == 1:                                 # convoy start: CTA-wide barrier
            if all(o['state'] >= 1 for o in cta['warps']):
                if not wp['active']:
                    wp['state'] = 6
                else:
                    wp['state'] = 2
                    heavy = rnd.random() < 0.03
                    wp['left'] = rnd.randint(40, 200) if heavy else rnd.randint(1, 6)
                return True
            return False
        if st == 2:                                 # the env step itself
            wp['left'] -= 1
            if wp['left'] <= 0:
                wp['state'] = 3
            return True
        if st == 3:                                 # publish: wait for the ring entry's counter to reach this launch's base
            if not wait_for_counter or signed(tails[cta['out']] - cta['base']) >= 0:
                wp['state'] = 4
                return True
            return False
        if st == 4:                                 # atomicAdd on the publish counter
            pos = (tails[cta['out']] - cta['base']) % MOD
            tails[cta['out']] = (tails[cta['out']] + 1) % MOD
            if not 0 <= pos < n:
                raise Violation(f"launch {cta['s']} publishes at position {pos} of a queue of {n}")
            wp['pos'] = pos
            wp['state'] = 5
            return True
        if st == 5:                                 # wait for the consumer of the slot's previous generation, then fill it
            want = ('E', (cta['gen_out'] - 1) % G)
            cur = q[cta['out']][wp['pos']]
            if cur == want or (not check_generation and cur[0] == 'E'):
                q[cta['out']][wp['pos']] = ('F', cta['gen_out'], wp['env'])
                next_launch[wp['env']] = cta['s'] + 1
                done_envsteps += 1
                wp['state'] = 6
                return True
            return False
        return False

    events = 0
    while done_envsteps < n * L:
        events += 1
        if events > max_events:
            raise Violation('event budget exhausted')
        try_place()
        live = [(cta, wp) for cta in resident for wp in cta['warps'] if wp['state'] < 6]
        rnd.shuffle(live)
        moved = False
        # a random subset of the resident warps makes one attempt each (spinning warps simply fail)
        for cta, wp in live[:max(1, len(live) // 3)]:
            moved |= step_warp(cta, wp)
        if not moved:                               # give every warp a chance before declaring a deadlock
            for cta, wp in live:
                moved |= step_warp(cta, wp)
        for cta in [c for c in resident if all(wp['state'] == 6 for wp in c['warps'])]:
            resident.remove(cta)
            free_sm += 1
            moved = True
        if not moved and not try_place():
            waiting = sorted({(c['s'], wp['state']) for c in resident for wp in c['warps'] if wp['state'] < 6})
            raise Violation(f'deadlock: resident launches / states {waiting[:12]}')
    # final state: the entry the next launch would read holds every env exactly once
    last = q[L % D]
    envs = sorted(v[2] for v in last if v[0] == 'F')
    if envs != list(range(n)) or any(x != L for x in next_launch):
        raise Violation('final queue does not hold every env once')
    return events


CASES = [  # n, W, D, S, L
    (7, 2, 2, 3, 40), (10, 3, 2, 8, 40), (12, 4, 3, 4, 30), (9, 3, 3, 16, 40), (16, 4, 8, 5, 60), (5, 1, 2, 6, 60),
    (20, 4, 4, 3, 30), (6, 2, 8, 12, 80), (11, 4, 3, 2, 30), (24, 8, 2, 4, 25), (3, 3, 2, 9, 60), (13, 2, 5, 7, 40),
]


@pytest.mark.parametrize('n,W,D,S,L', CASES)
def test_protocol_is_live_and_in_order(n, W, D, S, L):
    for seed in range(12):
        simulate(n, W, D, S, L, seed)


def test_small_moduli_wrap_many_times():
    """Counter modulus 256 and 4 generation tags: both wrap dozens of times in 120 launches (the real ones: 2^32 and 2048)."""
    for seed in range(8):
        simulate(6, 2, 2, 3, 120, seed, MOD=256, G=4)
        simulate(9, 3, 3, 4, 120, seed, MOD=256, G=8)


def _finds_violation(**kw):
    hits = 0
    for n, W, D, S, L in CASES:
        for seed in range(12):
            try:
                simulate(n, W, D, S, L, seed, max_events=300_000, **kw)
            except Violation:
                hits += 1
    return hits


def test_model_sees_the_missing_generation_tag():
    """The defect fixed by the slot generations: launches s and s + D spin on the same ring entry and the younger one takes the env."""
    assert _finds_violation(check_generation=False) > 0


def test_model_sees_the_missing_counter_wait():
    """The defect fixed by the publish-counter wait: a launch takes a position relative to the older launch that shares the counter."""
    assert _finds_violation(wait_for_counter=False) > 0
