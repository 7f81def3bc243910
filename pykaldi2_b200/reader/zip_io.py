"""Waveform / label ingestion from the reference's on-disk formats.

Formats kept as they are (reference reader/zip_io.py:129-160, reader/stream.py:519-575,
example/librispeech/README.md:20-45):
  * a corpus is a ``.zip`` of ``.wav`` files (or a directory tree); a member is addressed ``<zip>@/<member>``;
    the utterance id is the member's base name without extension (reader/stream.py:56-61);
  * label files are plain text, one utterance per line: ``utt-id int int int ...`` (pdf-ids for CE,
    transition-ids for the sequence losses); utterances without a label line in every label file are dropped.

Not a port: RIFF/WAVE is parsed here directly (the reference goes through ``soundfile``, which this image does not
have) into float32 in [-1, 1) exactly as soundfile's default float read scales integer PCM (x / 2^(bits-1)); the
waveforms go to the GPU fbank as they are, no feature extraction happens on the host.  FLAC members are
rejected with a clear error (no decoder in the image); a sample rate other than the target is resampled with
scipy's polyphase filter (the reference uses resampy: not bit-identical, LibriSpeech never takes this branch).
"""
import os
import struct
import zipfile

import numpy as np

SEP = "@/"


def parse_wav(blob, dtype=np.float32):
    """RIFF/WAVE bytes -> (sample_rate, samples float [n] or [n, channels])."""
    if len(blob) < 12 or blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(blob):
        cid, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        body = blob[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, fs, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == 0xFFFE and len(body) >= 26:            # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the tag
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, ch, fs, bits)
        elif cid == b"data":
            data = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError("WAVE file without fmt/data chunk")
    tag, ch, fs, bits = fmt
    if tag == 1:                                             # integer PCM
        if bits == 16:
            x = np.frombuffer(data, "<i2").astype(np.float64) / 32768.0
        elif bits == 8:
            x = (np.frombuffer(data, np.uint8).astype(np.float64) - 128.0) / 128.0
        elif bits == 32:
            x = np.frombuffer(data, "<i4").astype(np.float64) / 2147483648.0
        elif bits == 24:
            b = np.frombuffer(data[:len(data) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / 8388608.0
        else:
            raise ValueError("unsupported PCM width %d" % bits)
    elif tag == 3:                                           # IEEE float
        x = np.frombuffer(data, "<f4" if bits == 32 else "<f8").astype(np.float64)
    else:
        raise ValueError("unsupported WAVE format tag %d" % tag)
    x = x[:len(x) // ch * ch]
    if ch > 1:
        x = x.reshape(-1, ch)
    return int(fs), x.astype(dtype)


class ZipWaveIO(object):
    """``read_wav`` / ``walk`` over plain paths and ``<zip>@/<member>`` addresses (reference ZipWaveIO's surface)."""

    def __init__(self, precision="float32", fs=16000):
        self.dtype = np.dtype(precision)
        self.fs = int(fs)
        self._zips = {}

    def _zip(self, path):
        key = (os.getpid(), path)          # per process: a forked DataLoader worker must not share the parent's handle
        z = self._zips.get(key)
        if z is None:
            z = self._zips[key] = zipfile.ZipFile(path, "r")
        return z

    def close(self):
        for z in self._zips.values():
            z.close()
        self._zips = {}

    def __getstate__(self):            # DataLoader workers re-open the archives
        d = dict(self.__dict__)
        d["_zips"] = {}
        return d

    def read_bytes(self, name):
        if SEP in name:
            zpath, member = name.split(SEP, 1)
            return self._zip(zpath).read(member)
        with open(name, "rb") as f:
            return f.read()

    def read_wav(self, name):
        if name.lower().endswith(".flac"):
            raise ValueError("%s: FLAC needs a decoder this image does not have; convert the corpus to wav" % name)
        fs, x = parse_wav(self.read_bytes(name), self.dtype)
        if fs != self.fs:
            from math import gcd
            from scipy.signal import resample_poly
            g = gcd(fs, self.fs)
            x = resample_poly(x.astype(np.float64), self.fs // g, fs // g, axis=0).astype(self.dtype)
            fs = self.fs
        return fs, x

    def walk(self, zip_or_dir, extensions=(".wav",)):
        if zip_or_dir[-4:].lower() == ".zip":
            for member in self._zip(zip_or_dir).namelist():
                if member.lower().endswith(tuple(extensions)):
                    yield zip_or_dir + SEP + member
        else:
            for root, _, files in os.walk(zip_or_dir):
                for f in sorted(files):
                    if f.lower().endswith(tuple(extensions)):
                        yield os.path.join(root, f)


def utt_id_of(name):
    """Base name without extension (reader/stream.py:56-61)."""
    return os.path.splitext(os.path.basename(name.split("\t")[0]))[0]


def read_labels(path, wanted=None):
    """``utt-id int int ...`` per line -> {utt_id: int32 array}; ``wanted``: only these ids."""
    out = {}
    with open(path) as f:
        for line in f:
            parts = line.split()
            if not parts or (wanted is not None and parts[0] not in wanted):
                continue
            out[parts[0]] = np.asarray(parts[1:], dtype=np.int32)
    return out


def write_wav(path_or_file, samples, fs=16000):
    """16-bit PCM writer (test fixtures, tools)."""
    x = np.clip(np.round(np.asarray(samples, np.float64) * 32768.0), -32768, 32767).astype("<i2")
    ch = 1 if x.ndim == 1 else x.shape[1]
    data = x.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, ch, fs, fs * ch * 2, ch * 2, 16)
    blob = hdr + b"data" + struct.pack("<I", len(data)) + data
    if hasattr(path_or_file, "write"):
        path_or_file.write(blob)
    else:
        with open(path_or_file, "wb") as f:
            f.write(blob)
    return blob
