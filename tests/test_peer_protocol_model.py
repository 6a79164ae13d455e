"""Model check of the record exchange over peer memory (csrc/peer.cu, the wait in
csrc/step.cu k_pick_scatter, the send in csrc/selfuse.cuh sel_tail, and the order in which
markers/mutualinformation/iterative.py issues them).

The model: W ranks run the SAME script of operations; every rank owns a buffer
slot[half][src] = (tag, payload); a send with tag t writes (t, payload) into slot[t & 1][me] of
every rank; a wait for tag t blocks until all W slots of half t & 1 carry exactly t (a slot that
carries a LATER tag means the record was overwritten: the GPU kernel would spin until its
time-out).  All interleavings of the ranks are explored.  The shipped rule — strictly
send, wait, send, ... with a wait in front of every step that sends, also the add / remove steps
that use no record — never loses a record; the two rules it replaced do (which is what makes the
test meaningful)."""
import itertools

import pytest


def _program(script, wait_on_add, share_after_add_first):
    """The (op, tag) sequence one rank executes for a script of 'pick' / 'add' steps, as
    _DeviceScores.fused_step + Engine.fused_step issue them.  tag counts the sends."""
    prog, tag, records_valid = [], 0, False
    for step in script:
        if step == "pick" and not records_valid:
            tag += 1
            prog.append(("send", tag))  # _share_best -> prb_peer_share
            records_valid = True
        if step == "pick" or wait_on_add:
            prog.append(("wait", tag, step == "pick"))  # k_pick_scatter
        tag += 1
        prog.append(("send", tag))      # tail of the finalize kernel
        if not share_after_add_first:
            records_valid = True
    return prog


def _explore(prog, W):
    """DFS over all interleavings.  Returns None when every schedule completes with every pick
    reading the records of its own exchange, else a description of the failure."""
    empty = tuple(tuple((0, None) for _ in range(W)) for _ in range(2))
    start = (tuple(0 for _ in range(W)), tuple(empty for _ in range(W)))
    seen, stack = {start}, [start]
    while stack:
        pcs, bufs = stack.pop()
        if all(pc == len(prog) for pc in pcs):
            continue
        moved = False
        for r in range(W):
            if pcs[r] == len(prog):
                continue
            op = prog[pcs[r]]
            nb = bufs
            if op[0] == "send":
                t = op[1]
                nb = tuple(
                    tuple(tuple((t, (r, t)) if (h == (t & 1) and s == r) else bufs[d][h][s]
                                for s in range(W)) for h in range(2))
                    for d in range(W))
            else:
                _, t, reads = op
                slots = bufs[r][t & 1]
                if any(s[0] > t for s in slots):
                    return "rank %d waits for tag %d, slot already holds %s" % (r, t, slots)
                if any(s[0] != t for s in slots):
                    continue  # blocked
                if reads and t > 0 and any(s[1] != (src, t) for src, s in enumerate(slots)):
                    return "rank %d read a wrong record at tag %d" % (r, t)
            moved = True
            npcs = tuple(pc + 1 if i == r else pc for i, pc in enumerate(pcs))
            st = (npcs, nb)
            if st not in seen:
                seen.add(st)
                stack.append(st)
        if not moved:
            return "deadlock at %s" % (pcs,)
    return None


SCRIPTS = [s for n in (1, 2, 3, 4, 5) for s in itertools.product(("pick", "add"), repeat=n)]


@pytest.mark.parametrize("W", [2, 3])
def test_shipped_rule_never_loses_a_record(W):
    scripts = SCRIPTS if W == 2 else [s for s in SCRIPTS if len(s) <= 4]
    for script in scripts:
        prog = _program(script, wait_on_add=True, share_after_add_first=False)
        kinds = [op[0] for op in prog]  # never two sends without a wait in between
        assert not any(a == b == "send" for a, b in zip(kinds, kinds[1:])), script
        assert _explore(prog, W) is None, script


def test_rules_that_were_replaced_fail():
    # add / remove steps that send without waiting: a rank that is one step ahead overwrites
    # the record of the pick its peer is still waiting for
    bad = [s for s in SCRIPTS
           if _explore(_program(s, wait_on_add=False, share_after_add_first=False), 2)]
    assert ("pick", "add") in bad and ("pick",) not in bad and ("pick", "pick", "pick") not in bad
    # waiting on add, but a second send (prb_peer_share) before the first pick after an add
    bad = [s for s in SCRIPTS
           if _explore(_program(s, wait_on_add=True, share_after_add_first=True), 2)]
    assert ("add", "pick") in bad, bad[:4]
