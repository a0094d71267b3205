"""CPU model of the pooled epilogue's window bookkeeping (svision_b200/csrc/layer_tc.cu, STG = 0, and
`pool_crosses` in kernels.h): a layer's rows are cut into 128-row chunks; every 3x3/2 pooling window must be
written exactly once as a FIRST part (by the chunk that holds its first position, to pool_out) and, exactly
when its positions straddle a chunk boundary, exactly once as a SECOND part (by the next chunk, to
pool_out2); the two parts together must cover all nine positions, and a chunk never lists more windows than
the kernel's list holds.  841 and 196 are coprime with / do not divide 128, so a few hundred images run
through every alignment of image and chunk.  (The kernel itself is checked against the oracles on the GPU;
this pins the scheme it implements.)"""
import numpy as np
import pytest

BLOCK_M = 128
POOL_MAX_WINDOWS = 96          # layer_tc.cu


def pool_crosses(img, py, px, pos_per_img, grid_w):
    r0 = img * pos_per_img + 2 * py * grid_w + 2 * px
    return (r0 & (BLOCK_M - 1)) + 2 * grid_w + 2 >= BLOCK_M


def chunk_windows(chunk, n_img, pos_per_img, gw, valid, pool):
    """What the epilogue of one chunk lists: [(r0, img, py, px, part)] -- the rule of layer_tc.cu."""
    out = []
    m_rows = n_img * pos_per_img
    for t in range(BLOCK_M):
        row = chunk * BLOCK_M + t
        if row >= m_rows:
            continue
        img, q = divmod(row, pos_per_img)
        gy, gx = divmod(q, gw)
        if gy >= valid or gx >= valid:
            continue
        for a in (0, 1):
            py = (gy >> 1) - a
            if (a == 0 and py >= pool) or (a == 1 and ((gy & 1) or gy < 2)):
                continue
            for b in (0, 1):
                px = (gx >> 1) - b
                if (b == 0 and px >= pool) or (b == 1 and ((gx & 1) or gx < 2)):
                    continue
                dy, dx = gy - 2 * py, gx - 2 * px
                r0 = t - dy * gw - dx
                if dy == 0 and dx == 0:
                    mine = True
                elif r0 >= 0:
                    mine = False
                else:
                    first = -1
                    for yy in range(3):
                        rr = r0 + yy * gw
                        if first < 0 and rr + 2 >= 0:
                            first = 0 if rr < 0 else rr
                    mine = first == t
                if mine:
                    out.append((r0, img, py, px, 1 if r0 < 0 else 0))
    return out


@pytest.mark.parametrize("name,pos_per_img,gw,valid,pool,n_img", [
    ("conv2", 841, 29, 27, 13, 140),        # every alignment of 841-row images against 128-row chunks
    ("conv5", 196, 14, 13, 6, 70),
])
def test_every_window_is_written_once_per_part_and_fully_covered(name, pos_per_img, gw, valid, pool, n_img):
    assert valid == 2 * pool + 1 and 2 * gw + 3 < BLOCK_M
    m_rows = n_img * pos_per_img
    n_chunks = -(-m_rows // BLOCK_M)
    parts = {}                                           # (img, py, px, part) -> (chunk, covered positions)
    worst = 0
    for c in range(n_chunks):
        ws = chunk_windows(c, n_img, pos_per_img, gw, valid, pool)
        worst = max(worst, len(ws))
        for r0, img, py, px, part in ws:
            key = (img, py, px, part)
            assert key not in parts, (name, key, "listed twice")
            members = {c * BLOCK_M + r0 + yy * gw + xx for yy in range(3) for xx in range(3)
                       if 0 <= r0 + yy * gw + xx < BLOCK_M}
            parts[key] = (c, members)
    assert worst <= POOL_MAX_WINDOWS, (name, worst)
    # host-side bound of launch_layer
    own = ((BLOCK_M // gw) // 2 + 2) * pool
    cut = (((2 * gw + 2) // gw) // 2 + 1) * pool
    assert worst <= own + cut <= POOL_MAX_WINDOWS
    for img in range(n_img):
        for py in range(pool):
            for px in range(pool):
                first = img * pos_per_img + 2 * py * gw + 2 * px
                want = {first + yy * gw + xx for yy in range(3) for xx in range(3)}
                assert (img, py, px, 0) in parts, (name, img, py, px)
                c0, got = parts[(img, py, px, 0)]
                assert c0 == first // BLOCK_M                        # written by the chunk of its first position
                crosses = pool_crosses(img, py, px, pos_per_img, gw)
                assert crosses == ((img, py, px, 1) in parts), (name, img, py, px)
                if crosses:
                    c1, more = parts[(img, py, px, 1)]
                    assert c1 == c0 + 1 and not (got & more)
                    got = got | more
                assert got == want, (name, img, py, px)
    assert len(parts) == n_img * pool * pool + sum(1 for k in parts if k[3] == 1)
