"""CPU model of the lane / slot schedules of csrc/hist.cu (make_lane_const + accumulate16, tail_accumulate, the LDS.128
tile reads of hist_root_kernel): enumerates every step and proves the properties the kernel design relies on --
(1) every ATOMS instruction touches 32 distinct shared-memory banks, (2) every (row, feature slot) of a unit is updated
exactly once, (3) the shared-memory tile reads are bank-conflict free, (4) a window cannot overflow int32.
Keep in sync with hist.cu."""
import itertools


def lane_role(lane):                   # unit = 16 rows x one 32-feature group: two lanes per row
    return dict(row=lane >> 1, rot=lane >> 1, half=lane & 1, colbyte=(lane & 1) * 16)


def step_target(role, jw, jb):
    """(word rotation by rot>>2, byte select (jb + rot&3)&3) -> which of the lane's 16 bytes, and which slot."""
    qw, qb = role["rot"] >> 2, role["rot"] & 3
    word = (jw + qw) & 3               # the register read at step jw holds original word (jw + qw) & 3
    byte = (jb + qb) & 3
    k = 4 * word + byte                # byte index inside the lane's 16 B chunk
    return k, 16 * role["half"] + k


def test_main_unit_is_bank_conflict_free_and_covers_the_unit():
    seen = {}
    for jw, jb in itertools.product(range(4), range(4)):
        banks = set()
        for lane in range(32):
            role = lane_role(lane)
            k, slot = step_target(role, jw, jb)
            bank = slot % 32                                   # plane [256 bins][32 slots] int32: bank == slot; plane offsets are multiples of 32 words
            assert bank not in banks, (jw, jb, lane)
            banks.add(bank)
            seen[(role["row"], slot)] = seen.get((role["row"], slot), 0) + 1
            assert role["colbyte"] + k == slot                 # byte k of the lane's chunk IS that slot's bin
        assert len(banks) == 32
    assert len(seen) == 16 * 32 and set(seen.values()) == {1}


def test_gather_kernel_row_contiguous_mapping_is_conflict_free():
    """hist_gather_kernel: 2*NG adjacent lanes fetch one row's 32*NG contiguous bytes (16 / 8 / 5 rows per instruction);
    lane (row q, chunk c) = group c >> 1, half c & 1, rotation NG * q + (c >> 1); NG = 3 idles lanes 30, 31 on rotation 15."""
    for NG in (1, 2, 3):
        LPR, RPI = 2 * NG, 32 // (2 * NG)
        seen = {}
        for jw, jb in itertools.product(range(4), range(4)):
            banks = set()
            for lane in range(32):
                q, c = lane // LPR, lane % LPR
                on = q < RPI
                role = dict(rot=(NG * q + (c >> 1)) if on else 15, half=c & 1)
                k, slot = step_target(role, jw, jb)
                bank = slot % 32                               # group plane offsets are multiples of 32 words
                assert bank not in banks, (NG, jw, jb, lane)
                banks.add(bank)
                if on:
                    key = (q, c >> 1, slot)
                    seen[key] = seen.get(key, 0) + 1
                    assert (c & 1) * 16 + k == slot            # byte k of chunk c is slot 16 * half + k of group c >> 1
            assert len(banks) == 32
        assert len(seen) == RPI * NG * 32 and set(seen.values()) == {1}


def test_replicated_tail_is_bank_conflict_free_and_covers_the_unit():
    for tw in (4, 8):
        trep = 32 // tw                                        # plane [bin][trep][tw] int32
        seen = {}
        for j in range(tw):
            banks = set()
            for lane in range(32):                             # unit = 32 rows, one per lane
                slot = (j + lane) & (tw - 1)
                rep = (lane // tw) % trep
                bank = (rep * tw + slot) % 32                  # bin * trep * tw is a multiple of 32 words
                assert bank not in banks, (tw, j, lane)
                banks.add(bank)
                seen[(lane, slot)] = seen.get((lane, slot), 0) + 1
        assert len(seen) == 32 * tw and set(seen.values()) == {1}


def test_tile_reads_are_bank_conflict_free():
    """LDS.128 is served in quarter-warp phases of 8 lanes x 16 B; a phase is conflict free iff its 16 B units differ mod 8."""
    for ngc in (1, 3):                                         # tile row = 32 * ngc bytes (ngc == 2 is 2-way conflicted: documented in DESIGN.md)
        for g in range(ngc):
            for phase in range(4):
                units = set()
                for lane in range(8 * phase, 8 * phase + 8):
                    addr = (lane >> 1) * 32 * ngc + g * 32 + (lane & 1) * 16
                    units.add((addr // 16) % 8)
                assert len(units) == 8, (ngc, g, phase)


def test_fixed_point_window_cannot_overflow_int32():
    SPILL = 1 << 24
    for GRAD_BITS, WINDOW in ((18, 8064), (21, 1008)):                 # engine.h: large matrices / matrices up to 2^20 rows
        HESS_BITS = GRAD_BITS + 1
        assert (SPILL - 1) + WINDOW * (1 << GRAD_BITS) < 2 ** 31       # signed gradient plane
        assert (SPILL - 1) + WINDOW * (1 << HESS_BITS) < 2 ** 32       # unsigned hessian plane
        for sup, nwarps in ((32, 8), (32, 24), (30, 24)):              # gather kernel: super-tiles of 32 / 30 rows per warp between checks
            assert WINDOW // (sup * nwarps) >= 1 and (WINDOW // (sup * nwarps)) * sup * nwarps <= WINDOW
        assert WINDOW // 64 >= 1 and (WINDOW // 64) * 64 <= WINDOW     # root kernel: whole 64-row tiles between checks
