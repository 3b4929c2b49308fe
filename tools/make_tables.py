#!/usr/bin/env python3
"""Convert the reference's two data tables on the hot path into flat binary tables.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tools/make_tables.py
Outputs (committed; they are data the simulated cell needs at run time):
    nb_iot_environment_b200/data/lut2_u16.bin   uint16[7][8][501] = BLER * 2^15 for
        I_tbs in (0,1,2,3,4,6,8) x N_rep in (1,..,128) x SNR -30.0..20.0 step 0.1
        (reference system/LUT2.csv, loaded by system/channel.py:61-79; all 28 056 values are
        exact multiples of 2^-15, asserted below, so the table is lossless)
    nb_iot_environment_b200/data/mcs_u8.bin     uint8[500][30][3] = (I_tbs, N_rep, N_ru) for
        loss 121.4..171.3 step 0.1 x the 30 distinct TBS values
        (reference system/loss_tbs_to_mcs.p, used by system/node_b.py:117-120,619-622)
"""
import csv, os, pickle, sys
import numpy as np

REF = os.environ.get("NBIOT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nb_iot_environment_b200", "data")
ITBS = [0, 1, 2, 3, 4, 6, 8]
NREP = [1, 2, 4, 8, 16, 32, 64, 128]
TBS = [16, 24, 32, 40, 56, 72, 88, 104, 120, 144, 152, 176, 208, 256, 328, 344, 392, 408, 424, 440,
       504, 536, 552, 568, 600, 680, 808, 1000, 1096, 1384]


def main():
    os.makedirs(OUT, exist_ok=True)
    lut = np.full((7, 8, 501), 0xFFFF, dtype=np.uint32)
    with open(os.path.join(REF, "system", "LUT2.csv")) as f:
        rd = csv.DictReader(f)
        n = 0
        for row in rd:
            snr, tbs, itbs, nr, bler = float(row["snr"]), int(row["tbs"]), int(row["itbs"]), int(row["nr"]), float(row["bler"])
            assert tbs == 256
            i = int(round(snr * 10)) + 300
            assert 0 <= i <= 500 and abs(snr * 10 - round(snr * 10)) < 1e-9
            q = bler * 32768.0
            assert q == int(q) and 0 <= q <= 32768, (snr, itbs, nr, bler)
            lut[ITBS.index(itbs), NREP.index(nr), i] = int(q)
            n += 1
    assert n == 28056 and (lut != 0xFFFF).all()
    lut.astype("<u2").tofile(os.path.join(OUT, "lut2_u16.bin"))

    d = pickle.load(open(os.path.join(REF, "system", "loss_tbs_to_mcs.p"), "rb"))
    mcs = np.zeros((500, 30, 3), dtype=np.uint8)
    assert len(d) == 15000
    for li in range(500):
        # the reference looks the table up with round(loss, 1), i.e. numpy's rint(x*10)/10
        key = np.float64(1214 + li) / 10.0
        for ti, tbs in enumerate(TBS):
            v = d[(key, tbs)]          # KeyError here would mean the index form is not equivalent
            assert v[0] in ITBS and v[1] in NREP and 1 <= v[2] <= 10
            mcs[li, ti] = v
    assert (np.float64(171.4), 16) not in d   # the clip admits a loss with no row
    mcs.tofile(os.path.join(OUT, "mcs_u8.bin"))
    print("wrote", OUT, lut.shape, mcs.shape)


if __name__ == "__main__":
    sys.exit(main())
