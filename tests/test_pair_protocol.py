"""Protocol model of the persistent CTA-pair likelihood kernel in its LOCAL orientation (`scdef_b200/csrc/lik_flat.cuh::k_lik_flat`,
barrier block and wrappers in `lik_tc_pair.cuh`), CPU only.  (The GLOBAL orientation shares the rate-GEMM half of the
protocol; its second GEMM and Q hand-off are per CTA — plain single-CTA full/empty pairs.)

The kernel is five kinds of agents (per CTA: a COPY thread, epilogue warps; in the leader the MMA issuer, in the peer the
relay thread; plus the hardware: TMA copies and the tensor pipe complete asynchronously) that hand buffers to each
other through mbarriers with phase parities.  What goes wrong in such code is not arithmetic but the protocol: a wait on
the wrong parity, a barrier whose arrival count does not match its users, a buffer refilled while it is still being
read — on the device that is a hang or silent corruption, and the round's GPU minutes are few.  This file restates the
protocol (every wait / arrive / commit / copy of the kernel, in the kernel's order, with the kernel's counts and parity
expressions) as Python coroutines, runs them under many random interleavings and random completion delays of the
asynchronous agents, and checks: every agent finishes (no deadlock), no barrier is arrived on more often than its count
per phase, and no buffer is written while an earlier reader is still in flight or read before its writer finished.

The model covers the kernel's whole-pair-tile epilogue (`wide=True`, the shipped form) and the two variants it was chosen
from in round 2 (`wide=False`: one 64-gene half per barrier round; `fine=True`: the Q hand-off per k-step — both measured
slower or equal on a B200, profiles/r02_first_call_summary.txt, and removed from the kernel).  If the kernel's protocol
changes, change it here first."""
import random

import pytest

E_WARPS = 4   # epilogue warps per CTA in the model (16 in the kernel: the counts below are written in terms of it)
NQ = 2        # groups of epilogue warps by the k-step of the second GEMM their Q columns feed (4 in the kernel)


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phase = name, count, count, 0, 0

    def ready(self, parity):          # mbarrier.try_wait.parity: the phase with this parity has completed
        return (self.phase & 1) != parity

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, f"{self.name}: more arrivals than its count in one phase"
        self.pending -= 1
        self.tx += tx
        self._maybe_flip()

    def complete_tx(self, n):
        self.tx -= n
        assert self.tx >= 0, self.name
        self._maybe_flip()


class Resource:
    """A buffer with in-flight accesses.  Asynchronous agents (TMA, tensor pipe) hold their access from ISSUE to
    completion; `unread` counts the consumers that still have to read the current contents before it may be replaced."""

    def __init__(self, name):
        self.name, self.readers, self.writers, self.twriters, self.unread = name, 0, 0, 0, 0

    def begin_read(self):
        assert self.writers == 0 and self.twriters == 0, f"{self.name}: read while a write is in flight"
        self.readers += 1

    def begin_write(self):          # TMA or generic-proxy stores: exclusive
        assert self.writers == 0 and self.twriters == 0 and self.readers == 0, \
            f"{self.name}: write while {self.readers} reads / {self.writers + self.twriters} writes are in flight"
        assert self.unread == 0, f"{self.name}: overwritten before {self.unread} consumers have read it"
        self.writers += 1

    def begin_tensor_write(self, fresh):   # MMAs accumulate in order into the same tile: several may be in flight
        assert self.writers == 0 and self.readers == 0, f"{self.name}: tensor-core write while it is being read / copied"
        if fresh:
            assert self.unread == 0, f"{self.name}: accumulator restarted before {self.unread} consumers have read it"
        self.twriters += 1

    def consume(self):              # an epilogue warp reads the finished contents
        assert self.writers == 0 and self.twriters == 0, f"{self.name}: read before its writer finished"
        assert self.unread > 0, f"{self.name}: read twice or never produced"
        self.unread -= 1


class Sim:
    def __init__(self, n_outer, n_pt, pt0, wide, seed, fine=False):
        self.rng = random.Random(seed)
        self.n_outer, self.n_pt, self.pt0, self.wide, self.fine = n_outer, n_pt, pt0, wide, fine
        self.async_ops = []      # [due_time, callback]
        self.now = 0
        self.bars = [dict(), dict()]
        for r in range(2):
            b = self.bars[r]
            for name, cnt in (("z_full", 1), ("pz_full", 1), ("z_empty", 1), ("q_full", 2 * E_WARPS), ("q_empty", 1)):
                b[name] = Barrier(f"{name}@{r}", cnt)
            for k in range(NQ):     # fine mode: one Q hand-off per k-step group
                b[f"q_full_k{k}"] = Barrier(f"q_full_k{k}@{r}", 2 * E_WARPS // NQ)
                b[f"q_empty_k{k}"] = Barrier(f"q_empty_k{k}@{r}", 1)
            for i in range(2):
                b[f"w_full{i}"] = Barrier(f"w_full{i}@{r}", 1)
                b[f"pw_full{i}"] = Barrier(f"pw_full{i}@{r}", 1)
                b[f"w_empty{i}"] = Barrier(f"w_empty{i}@{r}", 1)
                b[f"r_full{i}"] = Barrier(f"r_full{i}@{r}", 1)
                b[f"r_empty{i}"] = Barrier(f"r_empty{i}@{r}", (2 if wide else 4) * E_WARPS)
                b[f"dz_full{i}"] = Barrier(f"dz_full{i}@{r}", 1)
                b[f"dz_empty{i}"] = Barrier(f"dz_empty{i}@{r}", 2 * E_WARPS)
        # buffers: per CTA z, W stage 0/1, Q; per pair (modelled once) R[2] and dz[2] in tensor memory
        self.res = {f"{n}@{r}": Resource(f"{n}@{r}") for r in range(2) for n in ["z", "w0", "w1", "q"] + [f"q_k{k}" for k in range(NQ)]}
        for n in ("R0", "R1", "dz0", "dz1"):
            self.res[n] = Resource(n)
        self.tensor_queue = []   # ops issued by the MMA thread, completed IN ORDER by the tensor pipe

    # ---- cursor over the flattened (sample, pair-tile) space -------------------------------------------------------
    def cursors(self):
        s, pt, g = 0, self.pt0, 0
        out = []
        for oi in range(self.n_outer):
            first = pt == 0 or oi == 0
            last = pt == self.n_pt - 1 or oi == self.n_outer - 1
            out.append((s, pt, g, first, last))
            pt += 1
            if pt == self.n_pt:
                pt, s, g = 0, s + 1, g + 1
        return out

    # ---- asynchronous agents ---------------------------------------------------------------------------------
    def later(self, fn):
        self.async_ops.append([self.now + self.rng.randint(1, 40), fn])

    def tma(self, rank, res_name, bar, nbytes):
        res = self.res[f"{res_name}@{rank}"]
        res.begin_write()

        def done():
            res.writers -= 1
            self.bars[rank][bar].complete_tx(nbytes)
        self.later(done)

    def mma(self, reads, write, fresh, unread_after=0):
        """Issue one group of pair MMAs (operands are held from issue to completion): `reads` are resource names (both CTAs'
        copies where per CTA), `write` the accumulator; fresh: accumulate flag 0; unread_after: consumers of the result."""
        rs = [self.res[n] for n in reads]
        w = self.res[write]
        for res in rs:
            res.begin_read()
        w.begin_tensor_write(fresh)
        self.tensor_queue.append(("mma", rs, w, unread_after))

    def commit(self, bar):      # multicast: the barrier at this offset in BOTH CTAs
        self.tensor_queue.append(("commit", bar))

    def tensor_pipe_step(self):
        if not self.tensor_queue:
            return False
        op = self.tensor_queue.pop(0)
        if op[0] == "commit":
            for r in range(2):
                self.bars[r][op[1]].arrive()
        else:
            for res in op[1]:
                res.readers -= 1
            op[2].twriters -= 1
            op[2].unread += op[3]
        return True

    # ---- the kernel's roles (generators: `yield (bar, parity)` = wait) ---------------------------------------------
    def copy_thread(self, rank):
        b = self.bars[rank]
        n_seg = 0
        for oi, (s, pt, g, first, last) in enumerate(self.cursors()):
            st = oi & 1
            yield b[f"w_empty{st}"], ((oi >> 1) & 1) ^ 1
            b[f"w_full{st}"].arrive(tx=7)
            for _ in range(7):                         # the stage's seven bulk copies
                pass
            self.tma(rank, f"w{st}", f"w_full{st}", 7)
            if first:
                yield b["z_empty"], (g & 1) ^ 1
                b["z_full"].arrive(tx=1)
                self.tma(rank, "z", "z_full", 1)
            n_seg = g + 1
        for st in range(2):
            uses = (self.n_outer + 1 - st) // 2
            if uses > 0:
                yield b[f"w_empty{st}"], (uses - 1) & 1
        if self.n_outer > 0:
            if self.fine:
                for k in range(NQ):
                    yield b[f"q_empty_k{k}"], (2 * self.n_outer - 1) & 1
            else:
                yield b["q_empty"], (2 * self.n_outer - 1) & 1
            yield b["z_empty"], (n_seg - 1) & 1

    def relay_thread(self):            # peer only: forwards "my TMA landed" to the leader
        b, lead = self.bars[1], self.bars[0]
        for oi, (s, pt, g, first, last) in enumerate(self.cursors()):
            yield b[f"w_full{oi & 1}"], (oi >> 1) & 1
            lead[f"pw_full{oi & 1}"].arrive()
            if first:
                yield b["z_full"], g & 1
                lead["pz_full"].arrive()

    def mma_thread(self):              # leader only
        b = self.bars[0]
        cur = self.cursors()
        U = 2 * self.n_outer
        next1 = 0

        def ready1(oi):
            s, pt, g, first, last = cur[oi]
            st, par = oi & 1, (oi >> 1) & 1
            if first and not (b["z_full"].ready(g & 1) and b["pz_full"].ready(g & 1)):
                return False
            return b[f"w_full{st}"].ready(par) and b[f"pw_full{st}"].ready(par) and b[f"r_empty{st}"].ready(par ^ 1)

        def issue1(oi):
            s, pt, g, first, last = cur[oi]
            st, par = oi & 1, (oi >> 1) & 1
            if first:
                yield b["z_full"], g & 1
                yield b["pz_full"], g & 1
            yield b[f"w_full{st}"], par
            yield b[f"pw_full{st}"], par
            yield b[f"r_empty{st}"], par ^ 1
            self.mma(["z@0", "z@1", f"w{st}@0", f"w{st}@1"], f"R{st}", fresh=True, unread_after=2 * 2 * E_WARPS)
            self.commit(f"r_full{st}")
            if last:
                self.commit("z_empty")

        for u in range(U):
            oi, h = u >> 1, u & 1
            st = oi & 1
            while True:
                if next1 < self.n_outer and next1 <= oi + 1 and (next1 == oi or ready1(next1)):
                    yield from issue1(next1)
                    next1 += 1
                    continue
                gate = b["q_full_k0"] if self.fine else b["q_full"]
                if gate.ready(u & 1):
                    break
                if next1 > oi + 1 or next1 >= self.n_outer:
                    yield gate, u & 1
                    break
                yield None                                  # poll again later
            s, pt, g, first, last = cur[oi]
            seg_first, seg_last = first and h == 0, last and h == 1
            dzb = g & 1
            if seg_first:
                yield b[f"dz_empty{dzb}"], ((g >> 1) & 1) ^ 1
            if self.fine:     # one k-step group at a time: its Q columns in, its MMAs, its slice handed back
                for k in range(NQ):
                    yield b[f"q_full_k{k}"], u & 1
                    self.mma([f"q_k{k}@0", f"q_k{k}@1", f"w{st}@0", f"w{st}@1"], f"dz{dzb}", fresh=seg_first and k == 0,
                             unread_after=2 * E_WARPS if (seg_last and k == NQ - 1) else 0)
                    self.commit(f"q_empty_k{k}")
            else:
                self.mma(["q@0", "q@1", f"w{st}@0", f"w{st}@1"], f"dz{dzb}", fresh=seg_first, unread_after=2 * E_WARPS if seg_last else 0)
                self.commit("q_empty")
            if h == 1:
                self.commit(f"w_empty{st}")
            if seg_last:
                self.commit(f"dz_full{dzb}")

    def epilogue_warp(self, rank, w):
        b, lead = self.bars[rank], self.bars[0]
        kq = w % NQ
        q = self.res[f"q_k{kq}@{rank}" if self.fine else f"q@{rank}"]
        q_full = lead[f"q_full_k{kq}" if self.fine else "q_full"]
        q_empty = b[f"q_empty_k{kq}" if self.fine else "q_empty"]
        cur = self.cursors()

        def store_q():
            q.begin_write()
            q.writers -= 1                      # generic-proxy stores complete before the fence + arrive
            q_full.arrive()

        def read_r(st):
            self.res[f"R{st}"].consume()

        def drain(g):
            dzb = g & 1
            yield b[f"dz_full{dzb}"], (g >> 1) & 1
            self.res[f"dz{dzb}"].consume()
            lead[f"dz_empty{dzb}"].arrive()

        if self.wide:
            for oi, (s, pt, g, first, last) in enumerate(cur):
                st, par = oi & 1, (oi >> 1) & 1
                yield b[f"w_full{st}"], par
                yield b[f"r_full{st}"], par
                read_r(st)                       # first half
                yield None
                read_r(st)                       # second half's tcgen05.ld, issued before the first Q store
                yield q_empty, 1
                store_q()
                lead[f"r_empty{st}"].arrive()    # both halves are in registers
                yield None
                yield q_empty, 0
                store_q()
                if last:
                    yield from drain(g)
        else:
            for u in range(2 * self.n_outer):
                oi, h = u >> 1, u & 1
                s, pt, g, first, last = cur[oi]
                st, par = oi & 1, (oi >> 1) & 1
                if h == 0:
                    yield b[f"w_full{st}"], par
                yield b[f"r_full{st}"], par
                read_r(st)
                lead[f"r_empty{st}"].arrive()
                yield None
                yield q_empty, (u & 1) ^ 1
                store_q()
                if h == 1 and last:
                    yield from drain(g)

    # ---- scheduler ---------------------------------------------------------------------------------------------------
    def run(self):
        agents = {}
        for r in range(2):
            agents[f"copy{r}"] = self.copy_thread(r)
            for w in range(E_WARPS):
                agents[f"epi{r}.{w}"] = self.epilogue_warp(r, w)
        agents["relay"] = self.relay_thread()
        agents["mma"] = self.mma_thread()
        waiting = {k: None for k in agents}          # current wait condition of each agent
        for k in list(agents):
            waiting[k] = self._advance(agents, k)
        steps = 0
        while agents:
            steps += 1
            assert steps < 2_000_000, "livelock"
            self.now += 1
            runnable = [k for k, cond in waiting.items() if k in agents and (cond is None or cond[0].ready(cond[1]))]
            due = [op for op in self.async_ops if op[0] <= self.now]
            choices = []
            if runnable:
                choices.append("agent")
            if due:
                choices.append("async")
            if self.tensor_queue and self.rng.random() < 0.5:
                choices.append("tensor")
            if not choices:
                if self.async_ops:
                    self.now = min(op[0] for op in self.async_ops)
                    continue
                if self.tensor_queue:
                    self.tensor_pipe_step()
                    continue
                stuck = {k: (c[0].name, c[1], c[0].phase, c[0].pending) for k, c in waiting.items() if k in agents and c}
                raise AssertionError(f"deadlock: {stuck}")
            pick = self.rng.choice(choices)
            if pick == "agent":
                k = self.rng.choice(runnable)
                waiting[k] = self._advance(agents, k)
            elif pick == "async":
                op = self.rng.choice(due)
                self.async_ops.remove(op)
                op[1]()
            else:
                self.tensor_pipe_step()
        while self.tensor_queue:
            self.tensor_pipe_step()
        for op in self.async_ops:
            op[1]()
        for r in range(2):
            for bar in self.bars[r].values():
                assert bar.tx == 0, bar.name

    def _advance(self, agents, k):
        try:
            return next(agents[k])
        except StopIteration:
            del agents[k]
            return None


@pytest.mark.parametrize("wide,fine", [(False, False), (True, False), (True, True), (False, True)])
@pytest.mark.parametrize("n_outer,n_pt,pt0", [(1, 1, 0), (2, 2, 0), (3, 2, 1), (5, 2, 0), (6, 3, 2), (7, 40, 38), (9, 4, 3), (8, 1, 0)])
def test_no_deadlock_no_overrun_no_hazard(wide, fine, n_outer, n_pt, pt0):
    """Chunks of n_outer pair-tiles starting at pair-tile pt0 of a sample with n_pt pair-tiles (so they cross 0..7 sample
    boundaries), 40 random schedules each.  fine: the Q hand-off per k-step group."""
    for seed in range(40):
        Sim(n_outer, n_pt, pt0, wide, seed, fine).run()


def _without_wait(role, prefix):
    """The model with one kind of wait removed from one role."""
    orig = getattr(Sim, role)

    def gen(self, *args):
        for cond in orig(self, *args):
            if cond is not None and cond[0].name.startswith(prefix):
                continue
            yield cond

    return type("Mutant", (Sim,), {role: gen})


@pytest.mark.parametrize("role,prefix", [
    ("copy_thread", "z_empty"), ("copy_thread", "w_empty"), ("epilogue_warp", "q_empty"), ("epilogue_warp", "r_full"),
    ("epilogue_warp", "dz_full"), ("mma_thread", "pw_full"), ("mma_thread", "pz_full"), ("mma_thread", "q_full"),
])
def test_the_model_catches_a_missing_wait(role, prefix):
    """Sanity of the checker: dropping any of these waits must end in a reported hazard, overrun or deadlock.  (Dropping
    the MMA thread's r_empty or dz_empty waits is NOT caught: in this protocol they are implied by the W-stage and Q
    hand-offs — the model says so, and that is worth knowing before spending a barrier on them.)"""
    mutant = _without_wait(role, prefix)
    with pytest.raises(AssertionError):
        for cfg in [(5, 2, 0), (6, 3, 2), (9, 4, 3), (8, 1, 0)]:
            for seed in range(30):
                mutant(*cfg, False, seed).run()


def test_the_model_catches_a_wrong_arrival_count():
    class Wrong(Sim):
        def __init__(self, *a):
            super().__init__(*a)
            for r in range(2):
                for i in range(2):
                    self.bars[r][f"r_empty{i}"] = Barrier(f"r_empty{i}@{r}", 4 * E_WARPS)   # the narrow epilogue's count

    with pytest.raises(AssertionError):
        for seed in range(10):
            Wrong(6, 3, 2, True, seed).run()


def test_the_model_catches_a_wrong_parity():
    """Sanity of the checker itself: the epilogue waiting on the wrong q_empty parity must be reported."""
    class Broken(Sim):
        def epilogue_warp(self, rank, w):
            gen = super().epilogue_warp(rank, w)
            for cond in gen:
                if cond is not None and cond[0].name.startswith("q_empty"):
                    cond = (cond[0], cond[1] ^ 1)
                yield cond

    with pytest.raises(AssertionError):
        for seed in range(20):
            Broken(4, 2, 0, False, seed).run()
