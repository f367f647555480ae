"""CPU model of the lane / slot schedule of csrc/hist.cu (accumulate16 + the LaneConst set-up): enumerates every step of
a tile and proves the two properties the kernel design relies on -- (1) every ATOMS instruction touches 32 distinct
shared-memory banks, (2) every (row, feature slot) of the tile is updated exactly once.  Keep in sync with hist.cu."""
import itertools

PLANE_WORDS = 256 * 32          # one int32 plane [256 bins][32 slots]; bank = word index % 32 = slot


def lane_roles(NG, lane):
    if NG == 2:                 # four lanes per row: 16 B chunks of the row's 64 B (group A low/high, group B low/high)
        q8, c = lane >> 2, lane & 3
        return dict(row=q8, rot=2 * q8 + (c >> 1), half=c & 1, group=c >> 1, colbyte=c * 16)
    return dict(row=lane >> 1, rot=lane >> 1, half=lane & 1, group=0, colbyte=(lane & 1) * 16)


def step_target(role, jw, jb):
    """(word rotation by rot>>2, byte select (jb + rot&3)&3) -> which of the lane's 16 bytes, and which slot."""
    qw, qb = role["rot"] >> 2, role["rot"] & 3
    word = (jw + qw) & 3        # the register read at step jw holds original word (jw + qw) & 3
    byte = (jb + qb) & 3
    k = 4 * word + byte         # byte index inside the lane's 16 B chunk
    slot = 16 * role["half"] + k
    return k, slot


def test_every_instruction_is_bank_conflict_free_and_covers_the_tile():
    for NG in (1, 2):
        rows_per_instr = 16 if NG == 1 else 8
        seen = {}
        for jw, jb in itertools.product(range(4), range(4)):
            banks = set()
            for lane in range(32):
                role = lane_roles(NG, lane)
                k, slot = step_target(role, jw, jb)
                bank = (role["group"] * 2 * PLANE_WORDS + slot) % 32       # plane offsets are multiples of 32 words
                assert bank not in banks, (NG, jw, jb, lane)
                banks.add(bank)
                key = (role["row"], role["group"], slot)
                seen[key] = seen.get(key, 0) + 1
                assert role["colbyte"] + k == (role["group"] * 32 + slot if NG == 2 else slot)   # byte k of the chunk IS that slot's bin
            assert len(banks) == 32
        assert len(seen) == rows_per_instr * NG * 32 and set(seen.values()) == {1}


def test_fixed_point_window_cannot_overflow_int32():
    GRAD_BITS, HESS_BITS, WINDOW, SPILL = 18, 19, 4096, 1 << 24       # engine.h / hist.cu constants
    assert (SPILL - 1) + WINDOW * (1 << GRAD_BITS) < 2 ** 31           # signed gradient plane
    assert (SPILL - 1) + WINDOW * (1 << HESS_BITS) < 2 ** 32           # unsigned hessian plane
    for nwarps in (8, 24, 32):
        assert (WINDOW // (32 * nwarps)) * 32 * nwarps <= WINDOW      # rows between two overflow checks
