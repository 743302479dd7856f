"""BASELINE.json configs[2] and configs[3] at their REAL shapes against the FP64 oracle (the whole-batch sizes --
3000 packets -- are covered by size-independent properties; the oracle comparison runs on full-shape packets).

  configs[2]  Nt32 Nr4, 1024 sc, SNR sweep -25..10 dB, FC 1024-1024-1024-1024 x2: 4 full packets per SNR level,
              tolerance per SNR for both tensor-core schemes (what tools/snr_sweep.py prints as a table)
  configs[3]  Nt64 Nr8, 2048 sc, FC 2048-1024-1024-2048 x2: 16 full packets (8192 pair rows), chunked workspace
  configs[2]/[3] batch: 3000-packet device-resident run == the same packets run in other chunkings / alone

Tolerances as the north_star states them: LS 1e-6, FC planes 1e-5 (rel-L2 vs the FP64 oracle).
"""
import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import tables, postproc
from _util import oracle_full, rel_l2, nmse_per_packet

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

TOL_LS = 1e-6
TOL_DNN = 1e-5


@pytest.mark.parametrize("precision", ["fp16x3", "tf32x3"])
def test_config3_snr_sweep_real_shape(precision):
    """configs[2] (SURVEY 8d "C3"): per-SNR tolerance at 32x4x1024 with the 1024-1024 nets, all 8 SNR levels."""
    nt, nr, nsc, hidden, npkt = 32, 4, 1024, (1024, 1024), 4
    x = mm.synth.make_pilots(nsc)
    nets = mm.synth.make_nets(nsc, hidden, nsc)
    P = tables.sylvester_hadamard(nt)
    table = []
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision=precision) as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        for si, snr in enumerate(range(-25, 11, 5)):
            Y, _ = mm.synth.make_packets(3, npkt, nt, nr, nsc, snr_db=float(snr), x_tones=x, first_pkt=npkt * si)
            Hr, Hi, Hls = eng.estimate(Y, want_ls=True)
            H, ref_r, ref_i = oracle_full(Y, P, x, 1, nets)
            e_ls = rel_l2(H, Hls)
            e_dnn = rel_l2(ref_r + 1j * ref_i, Hr.astype(np.float64) + 1j * Hi)
            n_dnn = nmse_per_packet(ref_r, ref_i, Hr, Hi, npkt, nr, nt)
            table.append((snr, e_ls, e_dnn, n_dnn))
    for snr, e_ls, e_dnn, n_dnn in table:
        assert e_ls <= TOL_LS, "SNR %d dB: H_ls %.2e" % (snr, e_ls)
        assert e_dnn <= TOL_DNN, "SNR %d dB: H_dnn %.2e" % (snr, e_dnn)
        assert n_dnn <= (2 * TOL_DNN) ** 2, "SNR %d dB: NMSE_subk %.2e" % (snr, n_dnn)


def test_config3_one_batch_holds_all_snr_levels():
    """The 3000-packet batch of configs[2] mixes SNR levels (noise-dominated packets at -25 dB are ~18x larger than the
    10 dB ones): one call over packets of all 8 levels, per-packet tolerance, device buffers."""
    import torch
    nt, nr, nsc, hidden = 32, 4, 1024, (1024, 1024)
    x = mm.synth.make_pilots(nsc)
    nets = mm.synth.make_nets(nsc, hidden, nsc)
    snrs = np.repeat(np.arange(-25.0, 11.0, 5.0), 2)
    Y, _ = mm.synth.make_packets(3, snrs.size, nt, nr, nsc, snr_db=snrs, x_tones=x, first_pkt=100)
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision="fp16x3") as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        Hr, Hi = eng.estimate(torch.from_numpy(Y).cuda())
        Hr, Hi = Hr.cpu().numpy(), Hi.cpu().numpy()
    _, ref_r, ref_i = oracle_full(Y, tables.sylvester_hadamard(nt), x, 1, nets)
    rows = nt * nr
    for p in range(snrs.size):
        sl = slice(p * rows, (p + 1) * rows)
        e = rel_l2(ref_r[sl] + 1j * ref_i[sl], Hr[sl].astype(np.float64) + 1j * Hi[sl])
        assert e <= TOL_DNN, "packet %d (SNR %g dB): %.2e" % (p, snrs[p], e)


@pytest.mark.parametrize("precision", ["fp16x3", "tf32x3"])
def test_config4_sixteen_packets_real_shape(precision):
    """configs[3] (SURVEY 8d "C4"): 64x8x2048, FC 2048-1024-1024-2048, 16 packets = 8192 pair rows, run through a
    6-packet workspace (chunks of 6 + 6 + 4) from host buffers."""
    nt, nr, nsc, hidden, npkt = 64, 8, 2048, (1024, 1024), 16
    x = mm.synth.make_pilots(nsc)
    nets = mm.synth.make_nets(nsc, hidden, nsc)
    Y, _ = mm.synth.make_packets(4, npkt, nt, nr, nsc, snr_db=10.0, x_tones=x)
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision=precision, max_pkts=6) as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        Hr, Hi, Hls = eng.estimate(Y, want_ls=True)
    ref_ls, ref_r, ref_i = oracle_full(Y, tables.sylvester_hadamard(nt), x, 1, nets)
    assert rel_l2(ref_ls, Hls) <= TOL_LS
    assert rel_l2(ref_r + 1j * ref_i, Hr.astype(np.float64) + 1j * Hi) <= TOL_DNN
    assert nmse_per_packet(ref_r, ref_i, Hr, Hi, npkt, nr, nt) <= (2 * TOL_DNN) ** 2
    r = mm.pair_row(11, 5, 40, nr, nt)                  # create_massiveMIMO_CSIest_dnn_dataset.py:62 at this shape
    assert rel_l2(ref_r[r], Hr[r]) <= 10 * TOL_DNN


def test_config3_full_batch_size_properties():
    """configs[2] at its full batch size (3000 packets, 384 000 pair rows, device-resident, fp16x3): the batch result
    of a tiled input is the tiled result, it equals the same packets pushed through other chunkings bitwise under a
    pinned operand scale, and three sampled packets agree with the oracle."""
    import torch
    nt, nr, nsc, hidden, npkt, distinct = 32, 4, 1024, (1024, 1024), 3000, 24
    x = mm.synth.make_pilots(nsc)
    nets = mm.synth.make_nets(nsc, hidden, nsc)
    snrs = np.tile(np.arange(-25.0, 11.0, 5.0), distinct // 8)
    Yg, _ = mm.synth.make_packets(3, distinct, nt, nr, nsc, snr_db=snrs, x_tones=x, first_pkt=500)
    Y = torch.from_numpy(np.concatenate([Yg] * (npkt // distinct))).cuda()          # 1.5 GiB of Y
    rows = nt * nr
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision="fp16x3", act_scale_log2=5) as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        Hr, Hi = eng.estimate(Y)                                  # 6 internal chunks of 512 packets
        torch.cuda.synchronize()
        blk = distinct * rows
        for rep in (1, 57, npkt // distinct - 1):
            assert torch.equal(Hr[:blk], Hr[rep * blk:(rep + 1) * blk]) and torch.equal(Hi[:blk], Hi[rep * blk:(rep + 1) * blk])
        h_small_r, h_small_i = eng.estimate(Y[:distinct])
        assert torch.equal(h_small_r, Hr[:blk]) and torch.equal(h_small_i, Hi[:blk])
        got_r, got_i = Hr[:blk].cpu().numpy(), Hi[:blk].cpu().numpy()
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision="fp16x3") as eng:           # automatic scale: same numbers to 1e-6
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        Ar, Ai = eng.estimate(Y)
        assert rel_l2(got_r, Ar[:blk].cpu().numpy()) <= 1e-6 and rel_l2(got_i, Ai[:blk].cpu().numpy()) <= 1e-6
    for p in (0, 9, 23):
        _, ref_r, ref_i = oracle_full(Yg[p:p + 1], tables.sylvester_hadamard(nt), x, 1, nets)
        sl = slice(p * rows, (p + 1) * rows)
        assert rel_l2(ref_r + 1j * ref_i, got_r[sl].astype(np.float64) + 1j * got_i[sl]) <= TOL_DNN, p
