#!/usr/bin/env python
"""Cross-entropy training of the BLSTM acoustic model on B200 (reference bin/train_ce.py).

Same flags and YAML schema as the reference script; ``-hvd`` selects multi-GPU (NCCL via torchrun
instead of Horovod).  ``-synthetic N`` trains on N seeded synthetic utterances; without it the
corpus of ``-data_config`` (zip of wavs + label files, the reference's formats) is read by
``data.SpeechDataset``.  The loop body is
bin/train_ce.py:177-208: features -> model -> CrossEntropyLoss(ignore_index=-100) -> backward ->
clip -> Adam(amsgrad) step, with fbank/CMN/chunking, the BLSTM and the loss on libpk2.so kernels.
"""
import argparse
import os
import pickle
import time

import numpy as np
import torch as th

import _common
from _common import pkdist
from pykaldi2_b200 import pipeline
from pykaldi2_b200.data.dataloader import SyntheticWaveDataset, WaveDataloader
from pykaldi2_b200.data.speech_dataset import SpeechDataset
from pykaldi2_b200.models import lstm
from pykaldi2_b200.reader.preprocess import GlobalMeanVarianceNormalization
from pykaldi2_b200.utils import utils


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-exp_dir")
    parser.add_argument("-dataPath", default='', type=str, help="path of data files")
    parser.add_argument("-train_config")
    parser.add_argument("-data_config")
    parser.add_argument("-lr", default=0.0001, type=float, help="Override the LR in the config")
    parser.add_argument("-batch_size", default=32, type=int, help="Override the batch size in the config")
    parser.add_argument("-data_loader_threads", default=0, type=int, help="number of workers for data loading")
    parser.add_argument("-max_grad_norm", default=5, type=float, help="max_grad_norm for gradient clipping")
    parser.add_argument("-sweep_size", default=200, type=float, help="process n hours of data per sweep (default:200)")
    parser.add_argument("-num_epochs", default=1, type=int, help="number of training epochs (default:1)")
    parser.add_argument("-global_mvn", default=False, type=_common.str2bool, help="if apply global mean and variance normalization")
    parser.add_argument("-resume_from_model", type=str, help="the model from which you want to resume training")
    parser.add_argument("-dropout", type=float, help="set the dropout ratio")
    parser.add_argument("-anneal_lr_epoch", default=2, type=int, help="start to anneal the learning rate from this epoch")
    parser.add_argument("-anneal_lr_ratio", default=0.5, type=float, help="the ratio to anneal the learning rate")
    parser.add_argument('-print_freq', default=100, type=int, metavar='N', help='print frequency (default: 100)')
    parser.add_argument('-hvd', default=False, type=_common.str2bool, help="multi-GPU data parallelism (NCCL; launch with torchrun)")
    parser.add_argument('-synthetic', default=0, type=int, help="train on this many seeded synthetic utterances")
    parser.add_argument('-seed', default=1234, type=int, help="random seed (model init, sampling)")
    parser.add_argument('-max_steps', default=0, type=int, help="stop after this many minibatches (0 = whole epoch)")
    parser.add_argument('-chunk_buffer', default=20000, type=int,
                        help="chunks kept in the shuffle buffer before minibatches are drawn at random from it "
                             "(the reference's DataBuffer holds 20000 samples)")
    args = parser.parse_args()

    th.manual_seed(args.seed)
    np.random.seed(args.seed)
    config = _common.load_config(args.train_config, args.data_config)
    config["sweep_size"] = args.sweep_size
    config["data_path"] = args.dataPath
    _common.dump_config(config)
    rank, world, local = _common.init_distributed(args.hvd)
    if not th.cuda.is_available():
        raise SystemExit("train_ce.py: the B200 build has no CPU path")
    dev = th.device("cuda", local)
    os.makedirs(args.exp_dir, exist_ok=True)

    mc, dc = config["model_config"], config["data_config"]
    if args.synthetic > 0:
        trainset = SyntheticWaveDataset(args.synthetic, mc["label_size"])
    else:                                   # the reference's corpus description: zip of wavs + label text files
        trainset = SpeechDataset(config)
    # the reference batches 80-frame chunks drawn at random from a sample buffer; here groups of utterances are cut
    # into chunks on the GPU and pooled in a device-side shuffle buffer (pipeline.ChunkPool)
    utts_per_batch = max(1, args.batch_size * dc.get("seg_len", 80) // 1230)
    loader = WaveDataloader(trainset, utts_per_batch, num_workers=args.data_loader_threads, distributed=world > 1)
    feat = pipeline.FeaturePipeline(use_cmn=dc.get("use_cmn", True))
    if args.global_mvn:
        transform = GlobalMeanVarianceNormalization()
        for i in range(min(len(trainset), 50)):
            wav = trainset[i][0]
            w, woff, foff = feat.ex.pack([wav])
            x, _ = feat.sequence_batch(w, woff, foff)
            transform.accumulate_stats(x[0].cpu().numpy())
        transform.learn_mean_and_variance_from_stats()
        feat.mvn = transform.device_vectors(dev)
        with open(args.exp_dir + "/transform.pkl", 'wb') as f:
            pickle.dump(transform, f, pickle.HIGHEST_PROTOCOL)
    print("Data loader set up successfully!")
    print("Number of minibatches: {}".format(len(loader)))

    dropout = mc["dropout"] if args.dropout is None else args.dropout
    model = lstm.LSTMAM(mc["feat_dim"], mc["label_size"], mc["hidden_size"], mc["num_layers"], dropout, True).to(dev)
    optimizer = th.optim.Adam(model.parameters(), lr=args.lr, amsgrad=True, fused=True)   # one launch for all parameters
    start_epoch = 0
    if args.resume_from_model:
        assert os.path.isfile(args.resume_from_model), "ERROR: model file {} does not exit!".format(args.resume_from_model)
        ckpt = _common.load_model_state(model, args.resume_from_model)
        start_epoch = ckpt.get("epoch", 0)
        optimizer.load_state_dict(ckpt["optimizer"])
        print("=> loaded checkpoint '{}' ".format(args.resume_from_model))
    if world > 1:
        pkdist.broadcast_parameters(model)
        pkdist.broadcast_optimizer_state(optimizer)
    averager = pkdist.GradAverager(list(model.parameters())) if world > 1 else None

    model.train()
    for epoch in range(start_epoch, args.num_epochs):
        if epoch > args.anneal_lr_epoch:
            for g in optimizer.param_groups:
                g['lr'] *= args.anneal_lr_ratio
        loader.set_epoch(epoch)
        if hasattr(trainset, "set_epoch"):
            trainset.set_epoch(epoch)               # chunk mode: this epoch's random sweep of the corpus
        run_train_epoch(model, optimizer, averager, feat, loader, epoch, args, dc)
        if rank == 0:
            _common.save_checkpoint(args.exp_dir + '/model.' + str(epoch) + '.tar', model, optimizer, epoch)


def run_train_epoch(model, optimizer, averager, feat, loader, epoch, args, dc):
    batch_time = utils.AverageMeter('Time', ':6.3f')
    losses = utils.AverageMeter('Loss', ':.4e')
    grad_norm = utils.AverageMeter('grad_norm', ':.4e')
    progress = utils.ProgressMeter(len(loader), batch_time, losses, grad_norm, prefix="Epoch: [{}]".format(epoch))
    rtf = utils.RTFMeter()
    seg_len, seg_shift = dc.get("seg_len", 80), dc.get("seg_shift", 80)
    dev = next(model.parameters()).device
    # room for the buffer plus the chunks of one more group of utterances (30 s utterance = 37 chunks)
    pool = pipeline.ChunkPool(args.chunk_buffer + args.batch_size + 64 * loader_group_size(loader), seg_len,
                              model.input_size, dev, seed=args.seed + epoch)
    state = {"step": 0, "end": time.time()}

    def train_step(x, y):
        prediction = model(x)
        loss = pipeline.ce_loss(prediction.view(-1, prediction.shape[2]), y.view(-1))
        loss.backward()
        norm = pipeline.finish_step(model, optimizer, averager, args.max_grad_norm)
        grad_norm.update(float(norm))
        losses.update(loss.item(), x.size(0))
        rtf.update(x.size(0) * seg_len)
        batch_time.update(time.time() - state["end"])
        state["end"] = time.time()
        if state["step"] % args.print_freq == 0:
            progress.print(state["step"])
            print("iRTF {:.1f}".format(rtf.irtf), flush=True)
        state["step"] += 1
        return bool(args.max_steps) and state["step"] >= args.max_steps

    done = False
    for data in loader:
        wav, woff, foff = feat.ex.pack(data["wav"])
        n_fr = [min(int(foff[u + 1] - foff[u]), len(l)) for u, l in enumerate(data["label"])]   # label trim
        x, cu, cs = feat.chunk_batch(wav, woff, foff, seg_len, seg_shift, n_fr)
        if x.shape[0] == 0:
            continue
        y = np.stack([data["label"][u][s:s + seg_len, 0] for u, s in zip(cu, cs)])
        pool.add(x, th.from_numpy(y).to(dev, non_blocking=True))
        # the reference pops samples once the buffer holds buffer_size of them (data/sr_dataset.py:70-73)
        while not done and pool.n >= max(args.chunk_buffer, args.batch_size):
            done = train_step(*pool.draw(args.batch_size))
        if done:
            return
    while not done and pool.n > 0:                  # end of the sweep: drain the buffer, nothing is discarded
        done = train_step(*pool.draw(args.batch_size))


def loader_group_size(loader):
    bs = getattr(loader, "batch_size", None)
    return int(bs) if bs else 64


if __name__ == '__main__':
    main()
