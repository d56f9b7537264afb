"""fvs_write_checkpoint: the device-resident hierarchy in the on-disk layout of
fub::amrex::WriteCheckpointFile (src/AMReX/PatchHierarchy.cpp:320-371: Header + one VisMF per
level).  AMReX is absent here, so the files are read back by a parser written from the
published format (Header: title, finest level, cycles, times, BoxArrays; VisMF header version 1:
how, ncomp, ngrow, BoxArray, FabOnDisk table, min / max tables; FAB: "FAB <RealDescriptor><box>
<ncomp>" line + native doubles, component slowest) and compared with the device state bit for
bit."""
import os
import re

import numpy as np
import pytest

from helpers import assert_bitwise_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fvs():
    from finitevolumesolver_b200 import capi
    assert capi.device_count() > 0, "no CUDA device"
    return capi


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


BOX = re.compile(r"\(\(([-\d,]+)\) \(([-\d,]+)\) \(([-\d,]+)\)\)")


def parse_boxarray(lines, pos):
    m = re.match(r"\((\d+) 0$", lines[pos])
    assert m, lines[pos]
    n = int(m.group(1))
    boxes = []
    for i in range(n):
        b = BOX.match(lines[pos + 1 + i])
        assert b, lines[pos + 1 + i]
        lo = [int(v) for v in b.group(1).split(",")]
        hi = [int(v) for v in b.group(2).split(",")]
        assert set(b.group(3).split(",")) == {"0"}  # cell centred
        boxes.append((lo, hi))
    assert lines[pos + 1 + n] == ")"
    return boxes, pos + 2 + n


def read_vismf(prefix):
    lines = open(prefix + "_H").read().split("\n")
    assert lines[0] == "1" and lines[1] == "1"      # version 1, NFiles
    ncomp = int(lines[2])
    assert set(lines[3].strip("()").split(",")) == {"0"}   # no ghost cells
    boxes, pos = parse_boxarray(lines, 4)
    assert int(lines[pos]) == len(boxes)
    fabs = []
    for b in range(len(boxes)):
        m = re.match(r"FabOnDisk: (\S+) (\d+)$", lines[pos + 1 + b])
        assert m, lines[pos + 1 + b]
        fabs.append((m.group(1), int(m.group(2))))
    pos += 1 + len(boxes)
    assert lines[pos] == ""
    tables = []
    for _ in range(2):
        assert lines[pos + 1] == f"{len(boxes)},{ncomp}"
        rows = [[float(v) for v in lines[pos + 2 + b].rstrip(",").split(",")] for b in range(len(boxes))]
        tables.append(np.array(rows))
        pos += 2 + len(boxes)
        assert lines[pos] == ""
    data = []
    for (lo, hi), (fname, offset) in zip(boxes, fabs):
        with open(os.path.join(os.path.dirname(prefix), fname), "rb") as f:
            f.seek(offset)
            head = f.readline().decode()
            m = re.match(r"FAB \(\(8, \(64 11 52 0 1 12 0 1023\)\),\(8, \(8 7 6 5 4 3 2 1\)\)\)(\(.*\)) (\d+)\n", head)
            assert m, head
            b = BOX.match(m.group(1))
            assert [int(v) for v in b.group(1).split(",")] == lo
            assert int(m.group(2)) == ncomp
            shape = [h - l + 1 for l, h in zip(lo, hi)][::-1]
            arr = np.frombuffer(f.read(8 * ncomp * int(np.prod(shape))), dtype="<f8")
            data.append(arr.reshape([ncomp] + shape))
    return boxes, data, tables


def test_checkpoint_round_trip(oracle, fvs, tmp_path):
    from test_gpu_amr import nested_hierarchy
    hier, cpus, gpus, ctx = nested_hierarchy(oracle, fvs, 3)
    for _ in range(3):
        dt = 0.5 * ctx.compute_stable_dt_hierarchy()
        ok, _, _ = ctx.advance_hierarchy(dt, 1)
        assert ok
    out = str(tmp_path / "chk000003")
    ctx.write_checkpoint(out)
    lines = open(os.path.join(out, "Header")).read().split("\n")
    assert lines[0] == "Checkpoint File" and lines[1] == "2"
    assert [int(v) for v in lines[2].split()] == [3, 6, 12]
    times = [float(v) for v in lines[3].split()]
    assert times == [ctx.time_point] * 3
    pos = 4
    g = 8
    for l in range(3):
        boxes, pos = parse_boxarray(lines, pos)
        assert [b[0] + [0] for b in boxes] == [list(b[0]) for b in gpus[l].boxes]
        vboxes, data, (mins, maxs) = read_vismf(os.path.join(out, f"Level_{l}", "data"))
        assert vboxes == boxes
        for b in range(len(boxes)):
            state = gpus[l].download_state(b)[:, :, g:-g, g:-g]
            assert_bitwise_equal(data[b][:, None], state, f"level {l} box {b}")
            assert np.array_equal(mins[b], state.reshape(state.shape[0], -1).min(axis=1))
            assert np.array_equal(maxs[b], state.reshape(state.shape[0], -1).max(axis=1))
    ctx.close()
