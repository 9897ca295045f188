"""Extract the reference's only numeric fixture (src/ExGrids.hs) into a compact .npz.

Run in the build container (the GPU box has no /root/reference):
    python tests/golden/make_exgrids_fixture.py
Fields follow src/ExGrids.hs:22-43: prevEvent (END 28 @ (2,4), t=1.7242386950809194),
prevGrid, prevFrep, prevAgent = (wNet, wGradCorr, [avgReward]), preCh = 30,
curEvent (NEW @ (3,3)), curGrid.  Only data is extracted; no reference code is copied.
"""
import os
import re

import numpy as np

SRC = "/root/reference/src/ExGrids.hs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exgrids.npz")


def int_list(text, name):
    m = re.search(r"^%s = \[([^\]]*)\]" % name, text, re.M)
    return np.array([int(v) for v in m.group(1).split(",")], dtype=np.int64)


def main():
    text = open(SRC).read()
    prev_grid = int_list(text, "prevGrid")
    prev_frep = int_list(text, "prevFrep")
    cur_grid = int_list(text, "curGrid")
    m = re.search(r"^prevAgent = \(\[([^\]]*)\],\s*\[([^\]]*)\],\s*\[([^\]]*)\]\)", text, re.M)
    wnet = np.array([float(v) for v in m.group(1).split(",")], dtype=np.float32)
    wgrad = np.array([float(v) for v in m.group(2).split(",")], dtype=np.float32)
    avgr = np.float32(float(m.group(3)))
    pre_ch = int(re.search(r"^preCh = (\d+)", text, re.M).group(1))
    ev = re.search(r"prevEvent = Event \{_evTime = ([0-9.e-]+), _evType = END (\d+) Nothing, _evCell = \((\d),(\d)\)\}", text)
    cur = re.search(r"curEvent = Event \{_evTime = ([0-9.e-]+), _evType = NEW, _evCell = \((\d),(\d)\)\}", text)
    assert prev_grid.size == 3430 and cur_grid.size == 3430 and prev_frep.size == 3479
    assert wnet.size == 3479 and wgrad.size == 3479
    np.savez_compressed(
        OUT,
        prev_grid=prev_grid.astype(np.uint8).reshape(7, 7, 70),
        prev_frep=prev_frep.astype(np.int16).reshape(7, 7, 71),
        cur_grid=cur_grid.astype(np.uint8).reshape(7, 7, 70),
        wnet=wnet, wgrad=wgrad, avg_reward=avgr,
        pre_ch=np.int64(pre_ch),
        prev_event_time=np.float64(ev.group(1)), prev_event_end_ch=np.int64(ev.group(2)),
        prev_event_cell=np.array([int(ev.group(3)), int(ev.group(4))]),
        cur_event_time=np.float64(cur.group(1)),
        cur_event_cell=np.array([int(cur.group(2)), int(cur.group(3))]),
    )
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
