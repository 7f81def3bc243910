#!/usr/bin/env python
"""Lattice-based sequence training (MMI / sMBR / MPFE) of the Transformer acoustic model on B200
(reference bin/train_transformer_se.py; SURVEY.md 8f-3, BASELINE config 5).

Same flags as the reference script (:54-82: -dim_model, -nheads, -ff_size, -nlayers, -look_ahead, -dropout,
-criterion, -ce_ratio ...).  Loop body = :222-291: time-major input, key-padding mask from num_frs, optional
look-ahead mask, frame-level CE (sum) for the CE-regularised objective, log-prior subtraction, the sequence
loss per utterance (batched by default here), SGD with momentum, clip, rank-0 checkpoints
``model.se.<epoch>.tar``.  Features come from the GPU fbank pipeline, lattices from ``SyntheticLatticeProvider``
(synthetic decoding lattices, as in train_se.py); gradients are averaged over ranks with NCCL (dist.py) instead
of Horovod.  The model runs on stock torch kernels under bf16 autocast (models/transformer.py).
"""
import argparse
import os
import time
import zlib

import numpy as np
import torch as th

import _common
from _common import pkdist
from pykaldi2_b200 import graphs, pipeline, synth
from pykaldi2_b200.data.dataloader import SyntheticWaveDataset, WaveDataloader
from pykaldi2_b200.data.speech_dataset import SpeechDataset
from pykaldi2_b200.models import transformer
from pykaldi2_b200.ops import ops
from pykaldi2_b200.utils import utils


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-config")
    parser.add_argument("-data", help="data yaml file")
    parser.add_argument("-dataPath", default='', type=str, help="path of data files")
    parser.add_argument("-seed_model", help="the seed nerual network model")
    parser.add_argument("-exp_dir", help="the directory to save the outputs")
    parser.add_argument("-transform", help="feature transformation matrix or mvn statistics")
    parser.add_argument("-criterion", type=str, choices=["mmi", "mpfe", "smbr"], default="mmi", help="set the sequence training crtierion")
    parser.add_argument("-trans_model", help="the HMM transistion model, used for lattice generation")
    parser.add_argument("-prior_path", help="the prior for decoder, usually named as final.occs in kaldi setup")
    parser.add_argument("-den_dir", help="the decoding graph directory to find HCLG and words.txt files")
    parser.add_argument("-lr", type=float, default=1e-5, help="set the learning rate")
    parser.add_argument("-ce_ratio", default=0.1, type=float, help="the ratio for ce regularization")
    parser.add_argument("-momentum", default=0, type=float, help="set the momentum")
    parser.add_argument("-batch_size", default=32, type=int, help="Override the batch size in the config")
    parser.add_argument("-dropout", default=0, type=float, help="set the dropout ratio")
    parser.add_argument("-nheads", default=4, type=int, help="the number of attention heads")
    parser.add_argument("-dim_model", default=512, type=int, help="the model dimension")
    parser.add_argument("-ff_size", default=2048, type=int, help="the size of feed-forward layer")
    parser.add_argument("-nlayers", default=6, type=int, help="the number of layers")
    parser.add_argument("-look_ahead", default=-1, type=int, help="the number of frames to look ahead")
    parser.add_argument("-data_loader_threads", default=0, type=int, help="number of workers for data loading")
    parser.add_argument("-max_grad_norm", default=5, type=float, help="max_grad_norm for gradient clipping")
    parser.add_argument("-sweep_size", default=100, type=float, help="process n hours of data per sweep (default:60)")
    parser.add_argument("-num_epochs", default=1, type=int, help="number of training epochs (default:1)")
    parser.add_argument('-print_freq', default=10, type=int, metavar='N', help='print frequency (default: 10)')
    parser.add_argument('-save_freq', default=1000, type=int, metavar='N', help='save model frequency (default: 1000)')
    parser.add_argument('-synthetic', default=0, type=int, help="train on this many seeded synthetic utterances")
    parser.add_argument('-silence_phones', default="1", type=str, help="colon separated silence phone ids (the reference reads <den_dir>/phones/silence.csl)")
    parser.add_argument('-batched_loss', default=1, type=int, help="0 = one loss call per utterance as the reference does")
    parser.add_argument('-seed', default=1234, type=int, help="random seed (model init, sampling)")
    parser.add_argument('-max_steps', default=0, type=int)
    args = parser.parse_args()

    th.manual_seed(args.seed)
    np.random.seed(args.seed)
    config = _common.load_config(args.config, args.data)
    config["sweep_size"] = args.sweep_size
    config["data_path"] = args.dataPath
    _common.dump_config(config)
    rank, world, local = _common.init_distributed(True)
    if not th.cuda.is_available():
        raise SystemExit("train_transformer_se.py: the B200 build has no CPU path")
    dev = th.device("cuda", local)
    os.makedirs(args.exp_dir, exist_ok=True)
    mc, dc = config["model_config"], config["data_config"]
    N = mc["label_size"]

    if args.synthetic > 0:
        dataset = SyntheticWaveDataset(args.synthetic, N)
    else:                                   # zip of wavs + pdf-id / transition-id label files (example/librispeech/README.md)
        dataset = SpeechDataset(config)
    loader = WaveDataloader(dataset, args.batch_size, num_workers=args.data_loader_threads, distributed=world > 1,
                            balanced=True, seed=args.seed)
    feat = pipeline.FeaturePipeline(use_cmn=dc.get("use_cmn", True))
    print("Data loader set up successfully!")
    print("Number of minibatches: {}".format(len(loader)))

    model = transformer.TransformerAM(mc["feat_dim"], args.dim_model, args.nheads, args.ff_size, args.nlayers,
                                      args.dropout, N).to(dev)
    optimizer = th.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum)
    if args.seed_model:
        _common.load_model_state(model, args.seed_model)
    if world > 1:
        pkdist.broadcast_parameters(model)
        pkdist.broadcast_optimizer_state(optimizer)
    averager = pkdist.GradAverager(list(model.parameters())) if world > 1 else None
    print(sum(int(np.prod(p.size())) for p in model.parameters() if p.requires_grad))

    rng = np.random.default_rng(1234)
    tid2pdf = np.concatenate([[-1], np.repeat(np.arange(N), 2)]).astype(np.int32)
    tid2phone = np.where(tid2pdf >= 0, tid2pdf // 3 + 1, 0).astype(np.int32)       # synthetic phones, 1 = silence
    trans_model = graphs.TidPdfMap(tid2pdf, tid2phone)
    if args.trans_model:                    # a Kaldi transition model in text form (copy-transition-model --binary=false)
        trans_model = graphs.TidPdfMap.from_kaldi_text(args.trans_model)
        if trans_model.num_pdfs() > N:
            raise SystemExit("%s: the transition model has %d pdfs, the network %d outputs" % (args.trans_model, trans_model.num_pdfs(), N))
    args.silence_ids = [int(i) for i in args.silence_phones.strip().split(':')]
    log_prior = th.from_numpy(synth.make_log_prior(N, rng)).to(dev)
    if args.prior_path:                     # final.occs in text form: log(occs / sum(occs)), bin/train_se.py:183-184
        from pykaldi2_b200.reader import kaldi_io
        lp = kaldi_io.log_prior_from_occs(args.prior_path)
        if len(lp) != N:
            raise SystemExit("%s: %d priors for %d network outputs" % (args.prior_path, len(lp), N))
        log_prior = th.from_numpy(lp).to(dev)
    asr_decoder = graphs.SyntheticLatticeProvider()
    if args.synthetic <= 0 and rank == 0:
        print("WARNING: this build has no decoder (SURVEY 8a row a12: HCLG decoding is outside the hot path): the "
              "denominator lattices are SYNTHETIC (random arcs around the alignment, seeded per utterance).  The loss "
              "kernels are exercised on the corpus' audio and alignments, but the model is not trained against real "
              "competing hypotheses.", flush=True)

    model.train()
    for epoch in range(args.num_epochs):
        loader.set_epoch(epoch)
        run_train_epoch(model, optimizer, averager, feat, log_prior, loader, epoch, asr_decoder, trans_model, args, rank)
        if rank == 0:
            _common.save_checkpoint(args.exp_dir + '/model.se.' + str(epoch) + '.tar', model, optimizer, epoch)


def run_train_epoch(model, optimizer, averager, feat, log_prior, loader, epoch, asr_decoder, trans_model, args, rank):
    batch_time = utils.AverageMeter('Time', ':6.3f')
    losses = utils.AverageMeter('Loss', ':.4e')
    grad_norm = utils.AverageMeter('grad_norm', ':.4e')
    progress = utils.ProgressMeter(len(loader), batch_time, losses, grad_norm, prefix="Epoch: [{}]".format(epoch))
    rtf = utils.RTFMeter()
    N = trans_model.num_pdfs()
    mmi = args.criterion == "mmi"
    end = time.time()
    for i, batch in enumerate(loader):
        wav, woff, foff = feat.ex.pack(batch["wav"])
        n_fr = [min(int(foff[u + 1] - foff[u]), len(l)) for u, l in enumerate(batch["label"])]   # data/sr_dataset.py:358-363
        x, num_frs = feat.sequence_batch(wav, woff, foff, n_frames=n_fr)            # [B, Tmax, F]
        B, Tmax = x.shape[0], x.shape[1]
        y = np.full((B, Tmax), -100, np.int64)
        for j, lab in enumerate(batch["label"]):
            y[j, :num_frs[j]] = lab[:num_frs[j], 0]
        y = th.from_numpy(y).cuda(non_blocking=True)
        # key-padding mask (True = padding) and look-ahead mask: bin/train_transformer_se.py:240-252
        frs = th.as_tensor(np.asarray(num_frs), device=x.device)
        key_padding_mask = th.arange(Tmax, device=x.device)[None, :] >= frs[:, None]
        src_mask = transformer.look_ahead_mask(Tmax, args.look_ahead, x.device) if args.look_ahead > -1 else None
        prediction = model(x.transpose(0, 1), src_mask, key_padding_mask).transpose(0, 1).contiguous()
        ce_loss = pipeline.ce_loss(prediction.view(-1, prediction.shape[-1]), y.view(-1), reduction="sum")
        lats, alis = [], []
        for j, ids in enumerate(batch["utt_ids"]):
            rng = np.random.default_rng(zlib.crc32(ids[0].encode()))
            trans_id = batch["aux"][j][0][0][:num_frs[j]].astype(np.int32)
            lat, _, _ = synth.make_lattice(int(num_frs[j]), N, rng, num_ali=trans_id,
                                               num_tids=trans_model.num_transition_ids())
            lats.append(graphs.Lattice(lat)); alis.append(trans_id)
        loglikes = prediction - log_prior
        if args.batched_loss:
            mpe = None if mmi else (args.criterion, trans_model.tid2phone, args.silence_ids)
            lb = graphs.LatticeBatch(lats, trans_model.tid2pdf, alis, device=x.device, mpe=mpe)
            se_loss = ops.MMIFunction.apply_batch(loglikes, lb) if mmi else ops.sMBRFunction.apply_batch(loglikes, lb)
        else:
            se_loss = 0.0
            for j in range(B):
                asr_decoder.push(lats[j])
                if mmi:
                    se_loss += ops.MMIFunction.apply(loglikes[j, :num_frs[j], :], asr_decoder, trans_model, alis[j].tolist())
                else:
                    se_loss += ops.sMBRFunction.apply(loglikes[j, :num_frs[j], :], asr_decoder, trans_model,
                                                      alis[j].tolist(), args.criterion, args.silence_ids)
        loss = se_loss.cuda() + args.ce_ratio * ce_loss
        loss.backward()
        norm = pipeline.finish_step(model, optimizer, averager, args.max_grad_norm)
        grad_norm.update(float(norm))
        tot_frs = np.array(num_frs).sum()
        losses.update(loss.item() / tot_frs)
        rtf.update(tot_frs)
        batch_time.update(time.time() - end)
        end = time.time()
        if rank == 0 and i > 0 and i % args.save_freq == 0:
            _common.save_checkpoint(args.exp_dir + '/model.se.' + str(epoch) + '.' + str(i) + '.tar', model, optimizer, epoch)
        if rank == 0 and i % args.print_freq == 0:
            progress.print(i)
            print("iRTF {:.1f}".format(rtf.irtf), flush=True)
        if args.max_steps and i + 1 >= args.max_steps:
            break


if __name__ == '__main__':
    main()
