"""Batch plans over device tensors (what bench.py drives) and host-buffer batch helpers.

torch is used for device memory and streams only; every kernel launch happens inside libb200z.so.
"""
import ctypes as C

import numpy as np

from . import _lib


class _Plan:
    def __init__(self, handle, n):
        self._h = handle
        self.n = n
        L = _lib.lib()
        self.in_bytes = L.b200z_plan_in_bytes(handle)
        self.out_bytes = L.b200z_plan_out_bytes(handle)
        self.in_offsets = np.array([L.b200z_plan_in_offset(handle, i) for i in range(n)], dtype=np.int64)
        # where slot i's data starts (behind a history / preset dictionary, if the plan has one)
        self.data_offsets = np.array([L.b200z_plan_data_offset(handle, i) for i in range(n)], dtype=np.int64)
        self.out_offsets = np.array([L.b200z_plan_out_offset(handle, i) for i in range(n)], dtype=np.int64)
        self.workspace_bytes = L.b200z_plan_workspace_bytes(handle)
        self.launches = L.b200z_plan_launches(handle)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().b200z_plan_destroy(h)
            except Exception:  # interpreter shutdown: the module globals may already be gone
                pass

    __del__ = close

    def set_timing(self, enable=True):
        _lib.raise_for(_lib.lib().b200z_plan_set_timing(self._h, 1 if enable else 0))

    def timings(self):
        """{kernel name: milliseconds} of the last run (call after synchronising the stream)."""
        names = C.create_string_buffer(1024)
        ms = (C.c_float * 32)()
        cnt = C.c_int32(0)
        _lib.raise_for(_lib.lib().b200z_plan_get_timings(self._h, names, 1024, ms, 32, C.byref(cnt)))
        keys = names.value.decode().split(";")[:cnt.value]
        return {k: float(ms[i]) for i, k in enumerate(keys)}

    def run(self, d_in, d_out, d_out_len, d_status, d_check=None, d_in_used=None, stream=None, stages=_lib.STAGE_ALL):
        """All arguments are CUDA torch tensors (uint8 blobs, int64 lengths, int32 status, uint32/int32 checks).
        Launches on `stream` (a torch.cuda.Stream) or the current stream; does not synchronise.
        stages: STAGE_SEARCH / STAGE_ENCODE run the two halves separately (b200z_plan_run_stages)."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        assert d_in.is_cuda and d_out.is_cuda and d_in.numel() >= self.in_bytes and d_out.numel() >= self.out_bytes
        rc = _lib.lib().b200z_plan_run_stages(self._h, d_in.data_ptr(), d_out.data_ptr(), d_out_len.data_ptr(),
                                              d_status.data_ptr(), d_check.data_ptr() if d_check is not None else None,
                                              d_in_used.data_ptr() if d_in_used is not None else None, stages, s.cuda_stream)
        _lib.raise_for(rc)


    def pack(self, d_out, d_out_len, d_packed, d_packed_off, stream=None):
        """streams of the last run back to back in d_packed; d_packed_off (int64, n + 1) gets starts and the total"""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        _lib.raise_for(_lib.lib().b200z_plan_pack(self._h, d_out.data_ptr(), d_out_len.data_ptr(), d_packed.data_ptr(),
                                                  d_packed_off.data_ptr(), s.cuda_stream))


class DeflatePlan(_Plan):
    """n independent streams, each what `new Deflater(level, true)` + SetInput(all) + Finish() would produce."""

    def __init__(self, in_lens, level=6, strategy=0, wrap=_lib.WRAP_RAW, end_mode=_lib.END_FINISH, dict_lens=None,
                 chunk_lens=None):
        """dict_lens: per-stream preset-dictionary bytes (Deflater.SetDictionary; at most 32506, the reference keeps the
        dictionary's tail) stored in the input slot directly in front of the stream's data.
        chunk_lens: per stream, the sizes of the SetInput calls that deliver its data (each followed by Deflate() until
        IsNeedingInput, as DeflaterOutputStream.Write does); levels 0-4 depend on it (SURVEY.md trap T9), 5-9 ignore it."""
        lens = np.ascontiguousarray(in_lens, dtype=np.int64)
        h = C.c_void_p()
        if chunk_lens is not None:
            assert len(chunk_lens) == lens.size
            arrs = [np.ascontiguousarray(c, dtype=np.int64) for c in chunk_lens]
            counts = np.array([a.size for a in arrs], dtype=np.int32)
            ptrs = (C.c_void_p * lens.size)(*[a.ctypes.data if a.size else None for a in arrs])
            dl = None if dict_lens is None else np.ascontiguousarray(dict_lens, dtype=np.int64)
            hs = _lib.History(_lib.HIST_NONE if dl is None else _lib.HIST_DICTIONARY, 0, None if dl is None else dl.ctypes.data,
                              None, None, None, counts.ctypes.data, C.addressof(ptrs), None, None, None)
            _lib.raise_for(_lib.lib().b200z_deflate_plan_create_ex(lens.size, lens.ctypes.data, level, strategy, wrap, end_mode,
                                                                  C.addressof(hs), C.byref(h)))
            self.dict_lens = np.zeros(lens.size, dtype=np.int64) if dl is None else dl
        elif dict_lens is None:
            _lib.raise_for(_lib.lib().b200z_deflate_plan_create(lens.size, lens.ctypes.data, level, strategy, wrap, end_mode,
                                                               C.byref(h)))
            self.dict_lens = np.zeros(lens.size, dtype=np.int64)
        else:
            dl = np.ascontiguousarray(dict_lens, dtype=np.int64)
            hs = _lib.History(_lib.HIST_DICTIONARY, 0, dl.ctypes.data, None, None, None)
            _lib.raise_for(_lib.lib().b200z_deflate_plan_create_ex(lens.size, lens.ctypes.data, level, strategy, wrap, end_mode,
                                                                  C.addressof(hs), C.byref(h)))
            self.dict_lens = dl
        self.in_lens = lens
        super().__init__(h, lens.size)


class InflatePlan(_Plan):
    """n independent raw deflate streams of comp_lens bytes decoding into at most out_caps bytes each."""

    def __init__(self, comp_lens, out_caps, wrap=_lib.WRAP_RAW, dict_lens=None):
        cl = np.ascontiguousarray(comp_lens, dtype=np.int64)
        oc = np.ascontiguousarray(out_caps, dtype=np.int64)
        h = C.c_void_p()
        dl = None if dict_lens is None else np.ascontiguousarray(dict_lens, dtype=np.int64)
        _lib.raise_for(_lib.lib().b200z_inflate_plan_create_ex(cl.size, cl.ctypes.data, oc.ctypes.data, wrap,
                                                              None if dl is None else dl.ctypes.data, C.byref(h)))
        self.dict_lens = dl
        self.comp_lens, self.out_caps = cl, oc
        super().__init__(h, cl.size)

    def set_start_bits(self, start_bits):
        """Raw plans: stream i's first block header starts at bit start_bits[i] (0..7) of its first compressed byte --
        decoding a stream on from a restart point, with the window image passed as the dictionary."""
        sb = np.ascontiguousarray(start_bits, dtype=np.int32)
        assert sb.size == self.n
        _lib.raise_for(_lib.lib().b200z_inflate_plan_set_start_bits(self._h, sb.ctypes.data))

    def stats(self, stream=None):
        """diagnostics of the last run (b200z_plan_get_stats): finder survivors, segments, round slots, blocks, rounds, passes,
        streams handed back to the serial kernel, pipeline in use"""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        v = np.zeros(8, dtype=np.uint32)
        _lib.raise_for(_lib.lib().b200z_plan_get_stats(self._h, v.ctypes.data, 8, s.cuda_stream))
        names = ("finder_survivors", "segments", "round_slots", "blocks", "rounds", "passes", "handed_back", "parallel")
        return dict(zip(names, (int(x) for x in v)))

    def restart_points(self, stream=None):
        """(bit, out_pos) int64 arrays of the last run: where the last block header each stream's decoder reached lies
        (bits from the first compressed byte of the slot; output bytes in front of it).  Synchronises the stream."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        bit = np.zeros(self.n, dtype=np.int64)
        pos = np.zeros(self.n, dtype=np.int64)
        _lib.raise_for(_lib.lib().b200z_plan_get_restart_points(self._h, bit.ctypes.data, pos.ctypes.data, s.cuda_stream))
        return bit, pos


class Pipeline:
    """Host-buffer pipeline (b200z_*_pipeline_*): submit() stages one batch from host memory and enqueues upload and kernels
    without waiting, collect() waits for the oldest batch and hands the produced bytes to host memory.  `ins` / `outs` are
    sequences of integer host addresses (ctypes / numpy / torch data pointers), one per stream."""

    def __init__(self, handle, n, depth):
        self._h, self.n, self.depth = handle, n, depth
        self.out_len = np.zeros(n, dtype=np.int64)
        self.in_used = np.zeros(n, dtype=np.int64)
        self.check = np.zeros(n, dtype=np.uint32)
        self.status = np.zeros(n, dtype=np.int32)

    @classmethod
    def deflate(cls, in_lens, level=6, strategy=0, wrap=_lib.WRAP_RAW, end_mode=_lib.END_FINISH, depth=2):
        lens = np.ascontiguousarray(in_lens, dtype=np.int64)
        h = C.c_void_p()
        _lib.raise_for(_lib.lib().b200z_deflate_pipeline_create(lens.size, lens.ctypes.data, level, strategy, wrap, end_mode, depth, C.byref(h)))
        return cls(h, lens.size, depth)

    @classmethod
    def inflate(cls, comp_lens, out_caps, wrap=_lib.WRAP_RAW, depth=2):
        cl = np.ascontiguousarray(comp_lens, dtype=np.int64)
        oc = np.ascontiguousarray(out_caps, dtype=np.int64)
        h = C.c_void_p()
        _lib.raise_for(_lib.lib().b200z_inflate_pipeline_create(cl.size, cl.ctypes.data, oc.ctypes.data, wrap, depth, C.byref(h)))
        return cls(h, cl.size, depth)

    @staticmethod
    def pointers(addresses):
        return (C.c_void_p * len(addresses))(*[int(a) if a else None for a in addresses])

    def submit(self, in_ptrs):
        _lib.raise_for(_lib.lib().b200z_pipeline_submit(self._h, in_ptrs))

    def collect(self, out_ptrs, out_caps, raise_on_error=True):
        rc = _lib.lib().b200z_pipeline_collect(self._h, out_ptrs, out_caps.ctypes.data, self.out_len.ctypes.data, self.in_used.ctypes.data,
                                              self.check.ctypes.data, self.status.ctypes.data)
        if raise_on_error or rc in (_lib.E_CUDA, _lib.E_ARG, _lib.E_STATE):
            _lib.raise_for(rc)
        return rc

    @property
    def in_flight(self):
        return int(_lib.lib().b200z_pipeline_in_flight(self._h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().b200z_pipeline_destroy(h)
            except Exception:
                pass

    __del__ = close


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data if a.size else None for a in arrs])


def deflate_batch(buffers, level=6, strategy=0, wrap=_lib.WRAP_RAW, end_mode=_lib.END_FINISH, devices=None):
    """Host buffers in, list of compressed bytes out (b200z_deflate_batch: H2D, kernels, D2H inside the call).
    devices: a list of CUDA device indices -- the batch is cut by bytes and every range runs on its own GPU at the same
    time (b200z_deflate_batch_multi, one host process for several GPUs).  Returns (outputs, checks)."""
    n = len(buffers)
    ins = [np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8)) for b in buffers]
    lens = np.array([a.size for a in ins], dtype=np.int64)
    caps = np.array([_lib.lib().b200z_deflate_bound(int(l)) + 16 for l in lens], dtype=np.int64)
    outs = [np.empty(int(c), dtype=np.uint8) for c in caps]
    out_len = np.zeros(n, dtype=np.int64)
    check = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    if devices is not None:
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        rc = _lib.lib().b200z_deflate_batch_multi(dv.ctypes.data, dv.size, _ptr_array(ins), lens.ctypes.data, n, level, strategy, wrap,
                                                  end_mode, _ptr_array(outs), caps.ctypes.data, out_len.ctypes.data,
                                                  check.ctypes.data, status.ctypes.data)
    else:
        rc = _lib.lib().b200z_deflate_batch(_ptr_array(ins), lens.ctypes.data, n, level, strategy, wrap, end_mode,
                                            _ptr_array(outs), caps.ctypes.data, out_len.ctypes.data, check.ctypes.data,
                                            status.ctypes.data)
    _lib.raise_for(rc)
    return [outs[i][:out_len[i]].tobytes() for i in range(n)], check


def inflate_batch(buffers, out_caps, raise_on_error=True, wrap=_lib.WRAP_RAW, return_checks=False, devices=None):
    """Host buffers in, (outputs, in_used, status) out.  status[i] = code | detail << 8.  wrap: WRAP_RAW (raw deflate),
    WRAP_ZLIB / WRAP_GZIP (header parsed, output checksum compared with the trailer on the device; gzip: one member,
    in_used says where the next one starts), WRAP_RAW_CRC32 (raw deflate, CRC-32 of the output reported: zip entries).
    return_checks: also return the checksums of the outputs."""
    n = len(buffers)
    ins = [np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8)) for b in buffers]
    lens = np.array([a.size for a in ins], dtype=np.int64)
    caps = np.ascontiguousarray(out_caps, dtype=np.int64)
    outs = [np.empty(int(c) + 1, dtype=np.uint8) for c in caps]
    out_len = np.zeros(n, dtype=np.int64)
    in_used = np.zeros(n, dtype=np.int64)
    check = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    if devices is not None:  # one host process, several GPUs (b200z_inflate_batch_multi)
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        rc = _lib.lib().b200z_inflate_batch_multi(dv.ctypes.data, dv.size, _ptr_array(ins), lens.ctypes.data, n, wrap, _ptr_array(outs),
                                                  caps.ctypes.data, out_len.ctypes.data, in_used.ctypes.data, check.ctypes.data,
                                                  status.ctypes.data)
    else:
        rc = _lib.lib().b200z_inflate_batch(_ptr_array(ins), lens.ctypes.data, n, wrap, _ptr_array(outs),
                                            caps.ctypes.data, out_len.ctypes.data, in_used.ctypes.data, check.ctypes.data,
                                            status.ctypes.data)
    if raise_on_error:
        _lib.raise_for(rc)
    elif rc in (_lib.E_CUDA, _lib.E_ARG):
        _lib.raise_for(rc)
    res = [outs[i][:out_len[i]].tobytes() for i in range(n)]
    if return_checks:
        return res, in_used, status, check
    return res, in_used, status


def zip_entries(buffers, level=6):
    """Compresses the payloads of many zip entries in one device batch.  Returns, per entry, what the reference's
    ZipOutputStream.PutNextPassthroughEntry (Zip/ZipOutputStream.cs:283-333) asks for: the raw deflate stream, the CRC-32
    and the size of the uncompressed bytes (CompressedSize is len(raw)) -- the container itself (local headers, central
    directory, Zip64, name encoding) stays with the reference's own ZipOutputStream."""
    outs, checks = deflate_batch(buffers, level=level, wrap=_lib.WRAP_RAW_CRC32)
    return [{"raw": o, "crc": int(c), "size": len(b)} for o, c, b in zip(outs, checks, buffers)]


def unzip_entries(raw_streams, sizes, crcs=None):
    """The read side (ZipFile.GetInputStream + the CRC test of ZipFile.TestArchive / ZipInputStream): inflates raw entry
    streams in one device batch, CRC-32 of every output computed on the device; raises on a mismatch with `crcs`."""
    outs, used, status, checks = inflate_batch(raw_streams, sizes, wrap=_lib.WRAP_RAW_CRC32, return_checks=True)
    if crcs is not None:
        for i, (got, want) in enumerate(zip(checks, crcs)):
            if int(got) != int(want) & 0xFFFFFFFF:
                raise _lib.SharpZipBaseException("CRC mismatch in entry %d" % i)
    return outs
