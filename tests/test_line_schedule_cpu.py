"""CPU: the scheduling algebra of the line-tile convolution kernel (bruno_b200/csrc/conv_sm100.cu: LineWalk, the issuer's
block / slot / accumulate rules, the commit rule), restated in Python and checked exhaustively on small geometries.

This is a model of the ALGORITHM, not of the CUDA code: it pins down the invariants the kernel relies on --
  * every output line (image group, y) receives exactly the contributions of the input lines y-1, y, y+1 that exist, in
    ascending input-line order (= the summation order dy = -1, 0, +1 of the per-tap kernel: the bit-identity argument);
  * the first contribution to an accumulator slot overwrites, the others accumulate;
  * a slot (ring of 8) is reused only after the line that owned it has been completed, and `tfull` is committed exactly
    once per output line, after its last contribution;
  * CTA ranges partition the output lines; a range that starts / ends inside an image reads one halo input line per side.
The GPU tests (tests/test_conv_gpu.py) check the kernel itself against float64 convolutions and, bit for bit, against the
per-tap kernel."""
import itertools

SLOTS = 8


def segments(H, t0, t1):
    """LineWalk: consecutive output lines yo0..yo1 (1-based) of one image group j, for the linear range [t0, t1)."""
    t = t0
    while t < t1:
        j = t // H
        yo0 = t - j * H + 1
        n = min(H - yo0 + 1, t1 - t)
        yield j, yo0, yo0 + n - 1
        t += n


def issue_plan(H, t0, t1):
    """What the MMA issuer does for the CTA range [t0, t1): a list of (group, input line yi, [(slot, output line, fresh)],
    [committed output lines])."""
    plan = []
    it_seg = 0
    for j, yo0, yo1 in segments(H, t0, t1):
        for yi in range(max(1, yo0 - 1), min(H, yo1 + 1) + 1):
            b_lo = 0 if yi - 1 >= yo0 else (1 if yi >= yo0 else 2)
            b_hi = 2 if yi + 1 <= yo1 else (1 if yi <= yo1 else 0)
            bf = 1 if yi == 1 else 2
            it0 = it_seg + (yi - 1 - yo0)
            blocks = [((it0 + b) % SLOTS, yi - 1 + b, b >= bf) for b in range(b_lo, b_hi + 1)]
            commits = []
            if b_lo == 0:
                commits.append((it0 % SLOTS, yi - 1))
            if yi == H and b_lo <= 1 <= b_hi:
                commits.append(((it0 + 1) % SLOTS, yi))
            plan.append((j, yi, blocks, commits))
        it_seg += yo1 - yo0 + 1
    return plan


def check_range(H, t0, t1):
    slot_owner = {}                      # slot -> (group, output line) currently accumulating there
    contrib = {}                         # (group, yo) -> list of input lines that contributed, in order
    done = []
    for j, yi, blocks, commits in issue_plan(H, t0, t1):
        for slot, yo, fresh in blocks:
            key = (j, yo)
            assert 1 <= yo <= H and abs(yo - yi) <= 1
            if fresh:
                assert key not in contrib, 'a line must be opened exactly once'
                assert slot not in slot_owner, 'slot %d reused while line %s is still accumulating' % (slot, slot_owner.get(slot))
                slot_owner[slot] = key
                contrib[key] = []
            else:
                assert slot_owner.get(slot) == key, 'accumulating into a slot that belongs to another line'
            contrib[key].append(yi)
        for slot, yo in commits:
            key = (j, yo)
            assert slot_owner.get(slot) == key
            del slot_owner[slot]         # the epilogue drains it; the gatekeeper waits for that before the slot is reopened
            done.append(key)
    assert not slot_owner, 'lines left uncommitted: %s' % slot_owner
    # every output line of the range exactly once, in order, with the right contributions in ascending order
    expect = [(t // H, t % H + 1) for t in range(t0, t1)]
    assert done == expect
    for (j, yo), ys in contrib.items():
        assert ys == [y for y in (yo - 1, yo, yo + 1) if 1 <= y <= H], ((j, yo), ys)
    return len(done)


def test_every_range_of_small_geometries():
    for H in (1, 2, 3, 5, 14):
        for groups in (1, 2, 3):
            n_lines = groups * H
            for t0, t1 in itertools.combinations(range(n_lines + 1), 2):
                assert check_range(H, t0, t1) == t1 - t0


def test_cta_partition_covers_every_line_once():
    """The kernel's range split: CTA c gets [c * n / grid, (c + 1) * n / grid)."""
    for H, groups, grid in ((28, 160, 148), (14, 80, 148), (32, 107, 148), (7, 5, 35), (3, 2, 6)):
        n = groups * H
        grid = min(grid, n)
        total = 0
        for c in range(grid):
            t0, t1 = c * n // grid, (c + 1) * n // grid
            total += check_range(H, t0, t1)
        assert total == n


def test_long_range_wraps_the_slot_ring_many_times():
    assert check_range(28, 3, 3 + 28 * 6 + 5) == 28 * 6 + 5
