"""Discrete-event model of the mbarrier protocol of ``tpconv_umma_kernel`` (csrc/tpconv_umma.cu), run on the CPU.

The kernel's roles (TMA producer, two MMA issuers, gather warp, two epilogue warpgroups -- one per accumulator) synchronise only through
mbarrier phase-parity waits, tcgen05.commit arrivals and an issue-order token.  Three bugs of that protocol were found
the hard way on the GPU (a token that was lapped when one issuer owned two slots in a row; a ring shorter than a tile's
slabs + 1 letting an issuer wait on a slot two fills behind; and, without the token, a parity test that aliases when bulk
copies land out of order -- test_no_token_aliases_when_bulk_copies_land_out_of_order), so this file restates the
protocol -- loop for loop, same parities, same barrier counts -- and runs it under many random interleavings,
checking what the hardware cannot tell us:

* no deadlock (every role finishes),
* no parity aliasing (a wait never passes before the phase it means, and is never lapped),
* no data hazard (slabs read before they landed / overwritten while read, accumulators overwritten before the
  epilogue read them or read before their MMAs completed, the A buffer overwritten while GEMMs still read it, GEMM2
  issued before the hidden activations exist).

It is a model, not the kernel: keep it in sync with the issue / epilogue / gather / producer loops when they change.
"""
import random

import pytest


class Barrier:
    """mbarrier: ``phase`` = index of the current (incomplete) phase; wait(parity) passes iff parity != phase & 1."""

    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f'{self.name}: more arrivals than the phase expects'
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passes(self, parity, intended):
        """``intended`` = the phase index the waiter means (model-only knowledge)."""
        ok = (self.phase & 1) != parity
        if ok:
            assert self.phase > intended, f'{self.name}: wait for phase {intended} passed in phase {self.phase} (parity aliasing)'
        assert self.phase <= intended + 1, f'{self.name}: waiter for phase {intended} lapped (barrier in phase {self.phase})'
        return ok


def acc_of(tt, nt):                       # tpconv_umma.cu: acc_of()
    return 0 if tt < 0 else (tt + nt) & 1


class Model:
    def __init__(self, items, ng, stages, nbuf, dual, rng, mutate=None, in_order_tma=False):
        self.items, self.ng, self.stages, self.nbuf, self.dual, self.rng = items, ng, stages, nbuf, dual, rng
        self.in_order_tma = in_order_tma
        self.mutate = mutate                  # None, or a deliberately broken / disabled piece of the protocol
        B = Barrier
        self.full = [B(f'full[{s}]', 1) for s in range(stages)]
        self.empty = [B(f'empty[{s}]', 1) for s in range(stages)]
        self.tmem_full = [B(f'tmem_full[{b}]', 1) for b in range(2)]
        self.tmem_empty = [B(f'tmem_empty[{b}]', 1) for b in range(2)]          # N_EPI arrivals modelled as one
        self.a_ready = [B(f'a_ready[{b}]', 1) for b in range(2)]
        self.a_free = [B(f'a_free[{b}]', 2 if dual else 1) for b in range(2)]
        self.h_ready = B('h_ready', 1)
        self.tok = [B('tok[0]', 1), B('tok[1]', 1)]
        # resources for the hazard checks
        self.slab = [None] * stages           # (item, tt, ks) currently valid in the slot, None = free / in flight
        self.slab_readers = [0] * stages      # MMA groups issued on the slot and not yet complete
        self.acc = [dict(state='free') for _ in range(2)]      # free -> writing(tag) -> done(tag) -> (epilogue reads) -> free
        self.abuf = [dict(content=None, readers=0) for _ in range(2)]   # ('X', item) / ('H', item)
        self.pipe = []                        # FIFO of issued MMA groups: dict(thread, stage, acc, tag, last, commits)
        self.fills = []                       # TMA copies in flight: (stage, tag)
        self.done = {}

    # ---------------------------------------------------------------- asynchronous hardware
    def async_events(self):
        ev = []
        if self.fills:
            ev.append('fill')
        if self.pipe:
            ev.append('mma')
        return ev

    def fire(self, kind):
        if kind == 'fill':
            # bulk copies that are in flight at the same time may land in ANY order (nothing orders their completion;
            # one slab can miss in L2 while a later one hits) unless the model is told otherwise
            stage, tag = self.fills.pop(0 if self.in_order_tma else self.rng.randrange(len(self.fills)))
            assert self.slab[stage] is None and self.slab_readers[stage] == 0, 'TMA wrote a slab that is still being read'
            self.slab[stage] = tag
            self.full[stage].arrive()
        else:
            g = self.pipe.pop(0)              # the tensor pipe completes MMA groups in issue order
            self.slab_readers[g['stage']] -= 1
            self.abuf[g['ab']]['readers'] -= 1
            for fn in g['commits']:
                fn()

    # ---------------------------------------------------------------- roles (generators: yield a predicate to wait on)
    def producer(self):
        stage, phase, fill_no = 0, 0, [0] * self.stages
        for it, nt in enumerate(self.items):
            for tt in range(-1, nt):
                for ks in range(self.ng):
                    k = fill_no[stage]
                    yield lambda s=stage, p=phase, k=k: self.empty[s].passes(p ^ 1, k - 1)
                    self.slab[stage] = None                      # the copy may land any time from now on
                    self.fills.append((stage, (it, tt, ks)))
                    fill_no[stage] += 1
                    stage += 1
                    if stage == self.stages:
                        stage, phase = 0, phase ^ 1
        self.done['producer'] = True

    def issuer(self, me):
        stage, phase, fill_no = 0, 0, [0] * self.stages
        te_phase, hr_phase, tok_phase, tok_pending = 0, 0, 0, 0
        te_n, tok_n = [0, 0], 0
        for it, nt in enumerate(self.items):
            ab = it % self.nbuf
            for tt in range(-1, nt):
                buf = acc_of(tt, nt)
                if self.dual and buf != me:
                    for _ in range(self.ng):
                        fill_no[stage] += 1
                        stage += 1
                        if stage == self.stages:
                            stage, phase = 0, phase ^ 1
                    tok_pending += 1
                    continue
                last_reader = tt + (2 if self.dual else 1) >= nt
                if tt < 0:
                    yield lambda ab=ab, it=it: self.a_ready[ab].passes((it // self.nbuf) & 1, it // self.nbuf)
                if tt == 0 or (self.dual and tt == 1):
                    yield lambda p=hr_phase, it=it: self.h_ready.passes(p, it)
                yield lambda b=buf, p=((te_phase >> buf) & 1) ^ 1, n=te_n[buf]: self.tmem_empty[b].passes(p, n - 1)
                te_phase ^= 1 << buf
                te_n[buf] += 1
                yield lambda s=stage, p=phase, k=fill_no[stage]: self.full[s].passes(p, k)
                if self.dual and tok_pending > 0 and self.mutate != 'no_token':
                    yield lambda p=tok_phase, n=tok_n: self.tok[me ^ 1].passes(p, n)
                    tok_phase ^= 1
                    tok_n += 1
                    tok_pending = 0
                # hazards at issue time
                assert self.acc[buf]['state'] == 'free', f'MMA into accumulator {buf} while it holds {self.acc[buf]}'
                self.acc[buf] = dict(state='writing', tag=(it, tt))
                want = ('X', it) if tt < 0 else ('H', it)
                for ks in range(self.ng):
                    if ks > 0:
                        yield lambda s=stage, p=phase, k=fill_no[stage]: self.full[s].passes(p, k)
                    assert self.slab[stage] == (it, tt, ks), f'MMA group reads slot {stage} holding {self.slab[stage]}, wants {(it, tt, ks)}'
                    assert self.abuf[ab]['content'] == want, f'MMA of {(it, tt)} reads A buffer holding {self.abuf[ab]["content"]}'
                    self.slab_readers[stage] += 1
                    self.abuf[ab]['readers'] += 1
                    commits = [lambda s=stage: (self._release(s), self.empty[s].arrive())]
                    if ks == self.ng - 1:
                        commits.append(lambda b=buf, tag=(it, tt): (self.acc.__setitem__(b, dict(state='done', tag=tag)), self.tmem_full[b].arrive()))
                        if last_reader:
                            commits.append(lambda ab=ab: self.a_free[ab].arrive())
                    self.pipe.append(dict(thread=me, stage=stage, ab=ab, commits=commits))
                    skip_token = tt < 0 and acc_of(0, nt) == 0 and self.mutate != 'token_every_tile'
                    if ks == self.ng - 1 and self.dual and not skip_token and self.mutate != 'no_token':
                        self.tok[me].arrive()
                    fill_no[stage] += 1
                    stage += 1
                    if stage == self.stages:
                        stage, phase = 0, phase ^ 1
                    yield lambda: True            # issue returns; other roles may run
            hr_phase ^= 1
        self.done[f'issuer{me}'] = True

    def _release(self, s):
        assert self.slab_readers[s] == 0, 'slot released with readers in flight'
        self.slab[s] = None

    def gather(self):
        for it, nt in enumerate(self.items):
            ab = it % self.nbuf
            yield lambda ab=ab, it=it: self.a_free[ab].passes(((it // self.nbuf) & 1) ^ 1, it // self.nbuf - 1)
            assert self.abuf[ab]['readers'] == 0, 'gather overwrites an A buffer that MMAs still read'
            self.abuf[ab]['content'] = ('X', it)
            self.a_ready[ab].arrive()
            yield lambda: True
        self.done['gather'] = True

    def epilogue(self, wg):
        """One epilogue warpgroup per accumulator: warpgroup ``wg`` consumes the tiles that land in accumulator ``wg``
        (acc_of); warpgroup 0 also converts the GEMM1 result at the start of every item (after its previous item's last
        own tile and flush -- no early conversion any more)."""
        tf_phase, tf_n = 0, 0
        for it, nt in enumerate(self.items):
            if wg == 0:
                yield lambda p=tf_phase, n=tf_n: self.tmem_full[0].passes(p, n)
                tf_phase ^= 1
                tf_n += 1
                assert self.acc[0] == dict(state='done', tag=(it, -1)), f'hidden activations of item {it} read from {self.acc[0]}'
                ab = it % self.nbuf
                assert self.abuf[ab]['readers'] == 0, 'hidden activations overwrite an A buffer that MMAs still read'
                self.abuf[ab]['content'] = ('H', it)
                self.acc[0] = dict(state='free')
                self.h_ready.arrive()
                self.tmem_empty[0].arrive()
            for tt in range(nt):
                if acc_of(tt, nt) != wg:
                    continue
                yield lambda p=tf_phase, n=tf_n: self.tmem_full[wg].passes(p, n)
                tf_phase ^= 1
                tf_n += 1
                assert self.acc[wg] == dict(state='done', tag=(it, tt)), f'tile {(it, tt)} read from {self.acc[wg]}'
                self.acc[wg] = dict(state='free')
                self.tmem_empty[wg].arrive()
                yield lambda: True
        self.done[f'epilogue{wg}'] = True

    # ---------------------------------------------------------------- scheduler
    def run(self):
        roles = {'producer': self.producer(), 'issuer0': self.issuer(0), 'gather': self.gather(), 'epilogue0': self.epilogue(0),
                 'epilogue1': self.epilogue(1)}
        if self.dual:
            roles['issuer1'] = self.issuer(1)
        waiting = {k: None for k in roles}
        steps = 0
        while roles:
            steps += 1
            assert steps < 2_000_000
            runnable = [k for k in roles if waiting[k] is None or waiting[k]()]
            choices = runnable + self.async_events()
            if not choices:
                raise AssertionError(f'deadlock: waiting roles {sorted(roles)}; phases ' +
                                     ', '.join(f'{b.name}={b.phase}' for b in self.tmem_full + self.tmem_empty + self.tok + [self.h_ready]))
            c = self.rng.choice(choices)
            if c in ('fill', 'mma'):
                self.fire(c)
                continue
            try:
                waiting[c] = next(roles[c])
            except StopIteration:
                del roles[c]
        while self.async_events():
            self.fire(self.async_events()[0])


CONFIGS = [
    # (NG, STAGES, NBUF, DUAL)                    kernel configuration
    (3, 4, 2, True),       # <60,10,64,false>: bf16 big model
    (6, 4, 1, False),      # <60,10,64,true>: bf16x3 big model (ring shorter than a tile + 1 -> single issuer)
    (3, 12, 2, True),      # KS = 32 bf16 (confidence / small models)
    (6, 11, 1, True),      # KS = 32 bf16x3: two issuers with a single A buffer
]


@pytest.mark.parametrize('ng,stages,nbuf,dual', CONFIGS)
def test_protocol_has_no_deadlock_aliasing_or_hazard(ng, stages, nbuf, dual):
    rng = random.Random(1234)
    for trial in range(60):
        n_items = rng.randint(1, 5)
        if trial % 3 == 0:
            items = [rng.choice([18, 23, 46])] * n_items                      # full items of one launch share n_tiles
        else:
            items = [rng.randint(1, 7) for _ in range(n_items)]                # tail-split parts: any length, mixed parity
        Model(items, ng, stages, nbuf, dual, random.Random(rng.random())).run()


def test_model_catches_the_short_ring_with_two_issuers():
    """The configuration the kernel refuses (DUAL needs STAGES >= NG + 1): the model must see it fail."""
    rng = random.Random(7)
    with pytest.raises(AssertionError):
        for _ in range(200):
            Model([rng.randint(2, 7) for _ in range(4)], 6, 4, 1, True, random.Random(rng.random())).run()


def test_model_catches_a_lapped_token():
    """Round-1 bug: an issuer that owns two slots in a row (GEMM1 + tile 0 of an even-length item) must signal the token
    once per run, not once per tile -- otherwise the other issuer, which waits once, is lapped."""
    rng = random.Random(3)
    with pytest.raises(AssertionError):
        for _ in range(200):
            Model([rng.choice([2, 4, 6])] * 3, 3, 4, 2, True, random.Random(rng.random()), mutate='token_every_tile').run()


def test_no_token_aliases_when_bulk_copies_land_out_of_order():
    """Root cause of the -DDDP_UMMA_TOKEN=0 failures on the GPU (cuda-gdb: "Warp Illegal Instruction" at the producer's
    ``mbarrier.arrive.expect_tx``, i.e. an arrival on a ``full`` barrier whose previous phase is still pending; only with two
    conv kernels in flight on two streams, profiles/r2_token0_rootcause.txt).  With a 4-slot ring and 3 slabs per tile
    an issuer that skips the other issuer's tile tests slot s for fill F while fill F-1 of s is one of the OTHER issuer's
    slabs.  Nothing orders the completion of bulk copies that are in flight together, so that slab can still be in the
    air when a later one has landed: the parity test then passes on the stale phase (aliasing), the MMAs read a slot
    that is being written, their commit frees it early and the producer re-arms a barrier that has not completed.
    With in-order completion the model (like most runs on the GPU) sees nothing; with out-of-order completion it must.
    The token closes the hole: it is released after the other issuer has ISSUED its tile, i.e. after it has seen all of that
    tile's slabs land, and it is awaited before the first non-blocking parity test of the own tile
    (test_protocol_has_no_deadlock_aliasing_or_hazard runs with out-of-order landing).  Letting the skipping issuer merely
    look at the foreign slabs' barriers instead does not work: it is not counted in ``empty``, so the slot can be recycled
    twice before it looks (lapped waiter)."""
    rng = random.Random(5)
    for _ in range(60):                                                       # in-order landing: consistent
        Model([rng.randint(1, 7) for _ in range(4)], 3, 4, 2, True, random.Random(rng.random()), mutate='no_token', in_order_tma=True).run()
    with pytest.raises(AssertionError, match='aliasing|holding|lapped|still being read|more arrivals'):
        for _ in range(400):
            Model([rng.randint(2, 7) for _ in range(4)], 3, 4, 2, True, random.Random(rng.random()), mutate='no_token').run()
