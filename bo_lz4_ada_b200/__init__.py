"""bo_lz4_ada_b200 -- B200-native LZ4 decompressor behind the LZ4Ada package API.

This package is a thin ctypes binding of the C-ABI in include/lz4b200.h
(bo_lz4_ada_b200/liblz4b200.so, built in-tree for sm_100a by csrc/Makefile).  It mirrors the
reference's Ada package `LZ4Ada` (/root/reference/lib/lz4ada.ads) name for name so that the
parity tests read like the reference's own test-suite:

    LZ4Ada.Init                -> Init(reservation)                    (lib/lz4ada.ads:189/218)
    LZ4Ada.Init_With_Header    -> Init_With_Header(data, reservation)  (:238)
    LZ4Ada.Init_For_Block      -> Init_For_Block(compressed_length)    (:255)
    Decompressor.Update        -> Decompressor.Update(data)            (:281)
    Is_End_Of_Frame            -> Decompressor.Is_End_Of_Frame()       (:303)
    XXHash32                   -> XXHash32.Hasher / XXHash32.Hash      (:311-321)
    exceptions                 -> Checksum_Error, Data_Corruption, Not_Supported,
                                  Too_Few_Header_Bytes, Too_Little_Memory (:133-162)
    (new) batched entry point  -> Batch / batch_decompress

There is no CPU decode path here: if the shared library or a CUDA device is missing, the calls
that need the GPU fail loudly (LibraryMissing / Device_Error).
"""
import ctypes
import os

__all__ = [
    "Init", "Init_With_Header", "Init_For_Block", "Decompressor", "XXHash32", "To_Hex",
    "LZ4AdaError", "Checksum_Error", "Data_Corruption", "Not_Supported", "Too_Few_Header_Bytes",
    "Too_Little_Memory", "Constraint_Error", "Assertion_Error", "Device_Error", "LibraryMissing",
    "DeviceContext", "Batch", "batch_decompress", "lib", "RESERVATIONS", "EOF_NAMES",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "liblz4b200.so")


class LibraryMissing(ImportError):
    pass


# Flexible_Memory_Reservation, lib/lz4ada.ads:79-106
RESERVATIONS = {"SZ_64_KiB": 0, "SZ_256_KiB": 1, "SZ_1_MiB": 2, "SZ_4_MiB": 3, "SZ_8_MiB": 4,
                "For_Modern": 3, "For_Legacy": 4, "For_All": 4, "Use_First": 5, "Single_Frame": 6}
EOF_NAMES = ["Yes", "No", "Maybe"]   # End_Of_Frame, lib/lz4ada.ads:124


class LZ4AdaError(Exception):
    """Base of the exceptions of lib/lz4ada.ads:133-162.  str(e) is the bare message,
    e.information the GNAT Exception_Information line ("raised LZ4ADA.X : message")."""
    ada_name = "?"

    def __init__(self, information):
        self.information = information
        msg = information.split(" : ", 1)[1] if " : " in information else information
        super().__init__(msg)


class Checksum_Error(LZ4AdaError):
    ada_name = "CHECKSUM_ERROR"


class Data_Corruption(LZ4AdaError):
    ada_name = "DATA_CORRUPTION"


class Not_Supported(LZ4AdaError):
    ada_name = "NOT_SUPPORTED"


class Too_Few_Header_Bytes(LZ4AdaError):
    ada_name = "TOO_FEW_HEADER_BYTES"


class Too_Little_Memory(LZ4AdaError):
    ada_name = "TOO_LITTLE_MEMORY"


class Constraint_Error(LZ4AdaError):
    ada_name = "CONSTRAINT_ERROR"


class Assertion_Error(LZ4AdaError):
    ada_name = "ASSERTION_ERROR"


class Device_Error(LZ4AdaError):
    """CUDA failure or no device: distinct from every LZ4 data error."""
    ada_name = "DEVICE_ERROR"


_EXC = [None, Checksum_Error, Data_Corruption, Not_Supported, Too_Few_Header_Bytes, Too_Little_Memory,
        Constraint_Error, Assertion_Error, Device_Error]
EXC_NAMES = ["OK"] + [c.ada_name for c in _EXC[1:]]

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_int_p = ctypes.POINTER(ctypes.c_int)


class BlkDesc(ctypes.Structure):      # lz4b200_blk_desc
    _fields_ = [("src_off", ctypes.c_uint64), ("dst_off", ctypes.c_uint64), ("src_len", ctypes.c_uint32),
                ("dst_cap", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("hist_avail", ctypes.c_uint32)]


class BlkStatus(ctypes.Structure):    # lz4b200_blk_status
    _fields_ = [("code", ctypes.c_uint32), ("out_len", ctypes.c_uint32), ("err_pos", ctypes.c_uint32),
                ("aux", ctypes.c_int32), ("xxh32_computed", ctypes.c_uint32), ("xxh32_declared", ctypes.c_uint32)]


class Chain(ctypes.Structure):        # lz4b200_chain
    _fields_ = [("first_block", ctypes.c_uint32), ("n_blocks", ctypes.c_uint32), ("dst_off", ctypes.c_uint64),
                ("dst_cap", ctypes.c_uint64)]


class HashSpan(ctypes.Structure):     # lz4b200_hash_span
    _fields_ = [("off", ctypes.c_uint64), ("len", ctypes.c_uint64)]


class XxhState(ctypes.Structure):     # lz4ada_xxhash32
    _fields_ = [("state", ctypes.c_uint32 * 4), ("buffer", ctypes.c_uint8 * 16), ("buffer_size", ctypes.c_int32),
                ("total_length", ctypes.c_uint64)]


class BatchItem(ctypes.Structure):    # lz4ada_batch_item
    _fields_ = [("src_off", ctypes.c_uint64), ("src_len", ctypes.c_uint64), ("dst_off", ctypes.c_uint64),
                ("dst_cap", ctypes.c_uint64)]


class BatchResult(ctypes.Structure):  # lz4ada_batch_result
    _fields_ = [("exception", ctypes.c_int32), ("end_of_frame", ctypes.c_int32), ("n_frames", ctypes.c_uint32),
                ("n_blocks", ctypes.c_uint32), ("dst_off", ctypes.c_uint64), ("out_len", ctypes.c_uint64)]


_SIGNATURES = {
    # device shim
    "lz4b200_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4b200_retain": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "lz4b200_sm_count": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_launch_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "lz4b200_set_tuning": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "lz4b200_get_tuning": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_k1_kernel_name": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_uint32]),
    "lz4b200_k1_fallbacks": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    "lz4b200_chain_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    "lz4b200_use_lane": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "lz4b200_sync_all": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4b200_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_alloc_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4b200_free_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_h2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "lz4b200_d2h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "lz4b200_memset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]),
    "lz4b200_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_timer_start": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_timer_stop": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "lz4b200_copy_stored": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_copy_probe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_float)]),
    "lz4b200_device_of": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_event_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4b200_event_destroy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_event_record": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_event_sync": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_event_elapsed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.POINTER(ctypes.c_float)]),
    "lz4b200_decode_blocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_decode_linked": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_xxh32_spans": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "lz4b200_xxh32_frames": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "lz4b200_size_blocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "lz4b200_stream_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4b200_stream_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_stream_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4b200_stream_block": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32,
                                            ctypes.POINTER(BlkStatus)]),
    "lz4b200_stream_block2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.POINTER(BlkStatus)]),
    "lz4b200_stream_digest": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32)]),
    "lz4b200_stream_adopt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]),
    "lz4b200_stream_adopt_list": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32),
                                                 ctypes.POINTER(ctypes.c_uint32), ctypes.c_int]),
    # LZ4Ada API
    "lz4ada_set_device_context": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4ada_init": (ctypes.c_int, [c_int_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4ada_init_with_header": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int_p, c_int_p, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]),
    "lz4ada_init_for_block": (ctypes.c_int, [c_int_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4ada_update": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_int_p, ctypes.c_void_p,
                                     ctypes.c_int, c_int_p, c_int_p]),
    "lz4ada_is_end_of_frame": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4ada_exception_message": (ctypes.c_char_p, [ctypes.c_void_p]),
    "lz4ada_free": (None, [ctypes.c_void_p]),
    "lz4ada_to_hex_u8": (None, [ctypes.c_uint8, ctypes.c_char_p]),
    "lz4ada_to_hex_u32": (None, [ctypes.c_uint32, ctypes.c_char_p]),
    "lz4ada_xxhash32_init": (None, [ctypes.POINTER(XxhState), ctypes.c_uint32]),
    "lz4ada_xxhash32_reset": (None, [ctypes.POINTER(XxhState), ctypes.c_uint32]),
    "lz4ada_xxhash32_update": (None, [ctypes.POINTER(XxhState), ctypes.c_void_p, ctypes.c_size_t]),
    "lz4ada_xxhash32_final": (ctypes.c_uint32, [ctypes.POINTER(XxhState)]),
    "lz4ada_xxhash32_hash": (ctypes.c_uint32, [ctypes.c_void_p, ctypes.c_size_t]),
    # batch
    "lz4ada_batch_plan": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32,
                                         ctypes.POINTER(BatchItem), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "lz4ada_batch_block_desc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(BlkDesc)]),
    "lz4ada_batch_host_outcome": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, c_int_p, c_int_p,
                                                 ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                                 ctypes.c_char_p, ctypes.c_size_t]),
    "lz4ada_batch_output_bytes": (ctypes.c_uint64, [ctypes.c_void_p]),
    "lz4ada_batch_block_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "lz4ada_batch_traffic": (None, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                    ctypes.POINTER(ctypes.c_uint64)]),
    "lz4ada_batch_kernel_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "lz4ada_batch_set_output_capacity": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64]),
    "lz4ada_batch_exact_sizing": (ctypes.c_int, [ctypes.c_void_p]),
    "lz4ada_batch_retried_streams": (ctypes.c_uint32, [ctypes.c_void_p]),
    "lz4ada_batch_k1_kernel_name": (ctypes.c_char_p, [ctypes.c_void_p]),
    "lz4ada_batch_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "lz4ada_batch_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "lz4ada_batch_run_pipelined": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_uint32]),
    "lz4ada_batch_results": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BatchResult)]),
    "lz4ada_last_k1_kernel_name": (ctypes.c_char_p, [ctypes.c_void_p]),
    "lz4ada_batch_decompress_multi": (ctypes.c_int, [ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_uint64,
                                                     ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(BatchItem),
                                                     ctypes.c_int, ctypes.POINTER(BatchResult), ctypes.c_char_p, ctypes.c_size_t]),
    "lz4ada_batch_message": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_uint32]),
    "lz4ada_batch_free": (None, [ctypes.c_void_p]),
    "lz4ada_batch_decompress": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                               ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(BatchItem),
                                               ctypes.c_int, ctypes.POINTER(BatchResult), ctypes.c_char_p,
                                               ctypes.c_size_t]),
}

_lib = None


def lib():
    """The loaded liblz4b200.so.  Raises LibraryMissing (never falls back to anything else)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LibraryMissing(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % SO_PATH)
        handle = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def _raise(kind, information):
    if isinstance(information, bytes):
        information = information.decode("utf-8", "replace")
    raise _EXC[kind](information)


def _as_buf(data):
    """bytes-like -> (ctypes address object, length) without copying where possible."""
    if isinstance(data, (bytes, bytearray)):
        n = len(data)
        if isinstance(data, bytes):
            return ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p), n
        return ctypes.cast((ctypes.c_uint8 * max(n, 1)).from_buffer(data), ctypes.c_void_p), n
    mv = memoryview(data).cast("B")
    n = len(mv)
    if mv.readonly:
        b = mv.tobytes()
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), n
    return ctypes.cast((ctypes.c_uint8 * max(n, 1)).from_buffer(mv), ctypes.c_void_p), n


def To_Hex(num, width=32):
    """To_Hex, lib/lz4ada.ads:306-307."""
    out = ctypes.create_string_buffer(9)
    if width == 8:
        lib().lz4ada_to_hex_u8(num, out)
    else:
        lib().lz4ada_to_hex_u32(num, out)
    return out.value.decode()


class DeviceContext:
    """lz4b200_ctx: one per GPU, used by one host thread at a time."""

    def __init__(self, device=0, stream=None):
        h = ctypes.c_void_p()
        rc = lib().lz4b200_create(device, ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if rc != 0:
            raise Device_Error("raised LZ4ADA.DEVICE_ERROR : lz4b200_create(device=%d) failed (rc=%d): no CUDA "
                               "device or driver. This library has no CPU decode path." % (device, rc))
        self.handle = h
        self.device = device

    def _ck(self, rc, what):
        if rc != 0:
            raise Device_Error("raised LZ4ADA.DEVICE_ERROR : %s failed: %s"
                               % (what, lib().lz4b200_last_error(self.handle).decode()))

    def set_tuning(self, blocks_per_warp):
        if lib().lz4b200_set_tuning(self.handle, blocks_per_warp) != 0:
            raise ValueError(blocks_per_warp)

    def k1_kernel_name(self, n_blocks):
        """Which K1 kernel lz4b200_decode_blocks launches for n_blocks under the current tuning."""
        return lib().lz4b200_k1_kernel_name(self.handle, int(n_blocks)).decode()

    def k1_fallbacks(self):
        """(blocks the last v6 launch handed to the exact routine, of those by the safety net)"""
        a, b = ctypes.c_uint32(0), ctypes.c_uint32(0)
        lib().lz4b200_k1_fallbacks(self.handle, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def chain_stats(self):
        """(blocks the chain kernel K7 finished, blocks it gave to the exact routine) since the last call"""
        a, b = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._ck(lib().lz4b200_chain_stats(self.handle, ctypes.byref(a), ctypes.byref(b)), "lz4b200_chain_stats")
        return a.value, b.value

    def alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._ck(lib().lz4b200_alloc(self.handle, nbytes, ctypes.byref(p)), "lz4b200_alloc")
        return p.value

    def free(self, ptr):
        self._ck(lib().lz4b200_free(self.handle, ctypes.c_void_p(ptr)), "lz4b200_free")

    def alloc_host(self, nbytes):
        p = ctypes.c_void_p()
        self._ck(lib().lz4b200_alloc_host(self.handle, nbytes, ctypes.byref(p)), "lz4b200_alloc_host")
        return p.value

    def free_host(self, ptr):
        self._ck(lib().lz4b200_free_host(self.handle, ctypes.c_void_p(ptr)), "lz4b200_free_host")

    def h2d(self, dst_dev, src, nbytes=None):
        if isinstance(src, int):
            addr = ctypes.c_void_p(src)
        else:
            addr, n = _as_buf(src)
            nbytes = n if nbytes is None else nbytes
        self._ck(lib().lz4b200_h2d(self.handle, ctypes.c_void_p(dst_dev), addr, nbytes), "lz4b200_h2d")
        self.sync()   # the source may be a temporary

    def d2h(self, src_dev, nbytes):
        out = bytearray(nbytes)
        if nbytes:
            addr, _ = _as_buf(out)
            self._ck(lib().lz4b200_d2h(self.handle, addr, ctypes.c_void_p(src_dev), nbytes), "lz4b200_d2h")
            self.sync()
        return bytes(out)

    def sync(self):
        self._ck(lib().lz4b200_sync(self.handle), "lz4b200_sync")

    def launch_count(self):
        return lib().lz4b200_launch_count(self.handle)

    def sm_count(self):
        return lib().lz4b200_sm_count(self.handle)

    def copy_probe(self, dst_dev, src_dev, nbytes, reps=5):
        """GB/s (read + write) of a plain device copy: the roofline denominator, measured in place."""
        ms = ctypes.c_float(0)
        self._ck(lib().lz4b200_copy_probe(self.handle, ctypes.c_void_p(dst_dev), ctypes.c_void_p(src_dev), nbytes, reps,
                                          ctypes.byref(ms)), "lz4b200_copy_probe")
        return 2.0 * nbytes / (ms.value / 1e3) / 1e9

    def make_default(self):
        lib().lz4ada_set_device_context(self.handle)

    def close(self):
        if self.handle:
            lib().lz4b200_destroy(self.handle)
            self.handle = None


class Decompressor:
    """LZ4Ada.Decompressor (lib/lz4ada.ads:126).  Owns the caller-side Buffer of Min_Buffer_Size
    bytes the way the reference's callers do (tool_unlz4ada_simple/unlz4ada_simple.adb:19)."""

    def __init__(self, handle, min_buffer_size):
        self._h = handle
        self.Min_Buffer_Size = min_buffer_size
        self.Buffer = bytearray(min_buffer_size)
        self._buf_c = (ctypes.c_uint8 * min_buffer_size).from_buffer(self.Buffer)

    def Update(self, data):
        """-> (Num_Consumed, output bytes, Output_First, Output_Last)   (lib/lz4ada.ads:281)"""
        addr, n = _as_buf(data)
        nc, of, ol = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        rc = lib().lz4ada_update(self._h, addr, n, ctypes.byref(nc), self._buf_c, self.Min_Buffer_Size,
                                 ctypes.byref(of), ctypes.byref(ol))
        if rc != 0:
            _raise(rc, lib().lz4ada_exception_message(self._h))
        out = bytes(self.Buffer[of.value:ol.value + 1]) if ol.value >= of.value else b""
        return nc.value, out, of.value, ol.value

    def Is_End_Of_Frame(self):
        return EOF_NAMES[lib().lz4ada_is_end_of_frame(self._h)]

    def close(self):
        if self._h:
            self._buf_c = None
            lib().lz4ada_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def Init(Reservation="For_All"):
    """LZ4Ada.Init (lib/lz4ada.ads:218) -> Decompressor (Min_Buffer_Size is an attribute)."""
    mb = ctypes.c_int(0)
    h = ctypes.c_void_p()
    rc = lib().lz4ada_init(ctypes.byref(mb), RESERVATIONS[Reservation], ctypes.byref(h))
    if rc != 0:
        _raise(rc, "raised ADA.ASSERTIONS.ASSERTION_ERROR : bad reservation")
    return Decompressor(h, mb.value)


def Init_With_Header(Input, Reservation="Single_Frame"):
    """LZ4Ada.Init_With_Header (lib/lz4ada.ads:238) -> (Decompressor, Num_Consumed)."""
    addr, n = _as_buf(Input)
    nc, mb = ctypes.c_int(0), ctypes.c_int(0)
    h = ctypes.c_void_p()
    msg = ctypes.create_string_buffer(700)
    rc = lib().lz4ada_init_with_header(addr, n, ctypes.byref(nc), ctypes.byref(mb), RESERVATIONS[Reservation],
                                       ctypes.byref(h), msg, 700)
    if rc != 0:
        _raise(rc, msg.value)
    return Decompressor(h, mb.value), nc.value


def Init_For_Block(Compressed_Length, Reservation="For_All"):
    """LZ4Ada.Init_For_Block (lib/lz4ada.ads:255)."""
    mb = ctypes.c_int(0)
    h = ctypes.c_void_p()
    rc = lib().lz4ada_init_for_block(ctypes.byref(mb), Compressed_Length, RESERVATIONS[Reservation], ctypes.byref(h))
    if rc != 0:
        _raise(rc, "raised ADA.ASSERTIONS.ASSERTION_ERROR : bad reservation")
    return Decompressor(h, mb.value)


class XXHash32:
    """package XXHash32 (lib/lz4ada.ads:311-321)."""

    class Hasher:
        def __init__(self, Seed=0):
            self._s = XxhState()
            lib().lz4ada_xxhash32_init(ctypes.byref(self._s), Seed)   # Init ignores Seed like the reference

        def Reset(self, Seed=0):
            lib().lz4ada_xxhash32_reset(ctypes.byref(self._s), Seed)

        def Update(self, Input):
            addr, n = _as_buf(Input)
            lib().lz4ada_xxhash32_update(ctypes.byref(self._s), addr, n)

        def Final(self):
            return lib().lz4ada_xxhash32_final(ctypes.byref(self._s))

    @staticmethod
    def Init(Seed=0):
        return XXHash32.Hasher(Seed)

    @staticmethod
    def Hash(Input):
        addr, n = _as_buf(Input)
        return lib().lz4ada_xxhash32_hash(addr, n)


class Batch:
    """The batched device entry point: plan (host) -> upload -> run (device) -> results."""

    def __init__(self, ctx, src, items, Reservation="For_All"):
        """src: bytes-like holding every stream; items: list of (src_off, src_len) or
        (src_off, src_len, dst_off, dst_cap)."""
        self.ctx = ctx
        self._src_keep = src
        self.src_addr, self.src_len = (ctypes.c_void_p(src), None) if isinstance(src, int) else _as_buf(src)
        self.n = len(items)
        arr = (BatchItem * max(self.n, 1))()
        for k, it in enumerate(items):
            arr[k].src_off, arr[k].src_len = it[0], it[1]
            if len(it) > 2:
                arr[k].dst_off, arr[k].dst_cap = it[2], it[3]
        self._items = arr
        h = ctypes.c_void_p()
        total = self.src_len if self.src_len is not None else max((i[0] + i[1]) for i in items)
        self.src_bytes = total
        rc = lib().lz4ada_batch_plan(ctx.handle if ctx is not None else None, self.src_addr, total, self.n, arr, RESERVATIONS[Reservation],
                                     ctypes.byref(h))
        if rc != 0:
            _raise(rc, "raised LZ4ADA.%s : lz4ada_batch_plan failed" % EXC_NAMES[rc])
        self._h = h

    @property
    def output_bytes(self):
        return lib().lz4ada_batch_output_bytes(self._h)

    @property
    def block_count(self):
        return lib().lz4ada_batch_block_count(self._h)

    def block_desc(self, index):
        d = BlkDesc()
        if lib().lz4ada_batch_block_desc(self._h, index, ctypes.byref(d)) != 0:
            raise IndexError(index)
        return d

    def host_outcome(self, item):
        """What the host stage alone concluded (no device needed)."""
        exc, eof = ctypes.c_int(0), ctypes.c_int(0)
        nf, nb = ctypes.c_uint32(0), ctypes.c_uint32(0)
        msg = ctypes.create_string_buffer(700)
        lib().lz4ada_batch_host_outcome(self._h, item, ctypes.byref(exc), ctypes.byref(eof), ctypes.byref(nf),
                                        ctypes.byref(nb), msg, 700)
        return {"exception": EXC_NAMES[exc.value], "end_of_frame": EOF_NAMES[eof.value], "n_frames": nf.value,
                "n_blocks": nb.value, "message": msg.value.decode()}

    def upload(self, src_dev, copy_source=True):
        rc = lib().lz4ada_batch_upload(self._h, self.src_addr if copy_source else None, ctypes.c_void_p(src_dev))
        if rc != 0:
            _raise(rc, "raised LZ4ADA.DEVICE_ERROR : lz4ada_batch_upload: %s"
                   % lib().lz4b200_last_error(self.ctx.handle).decode())

    def run(self, src_dev, dst_dev):
        rc = lib().lz4ada_batch_run(self._h, ctypes.c_void_p(src_dev), ctypes.c_void_p(dst_dev))
        if rc != 0:
            _raise(rc, "raised LZ4ADA.%s : lz4ada_batch_run: %s"
                   % (EXC_NAMES[rc], lib().lz4b200_last_error(self.ctx.handle).decode()))

    def results(self):
        res = (BatchResult * max(self.n, 1))()
        lib().lz4ada_batch_results(self._h, res)
        out = []
        for k in range(self.n):
            r = res[k]
            out.append({"exception": EXC_NAMES[r.exception], "end_of_frame": EOF_NAMES[r.end_of_frame],
                        "n_frames": r.n_frames, "n_blocks": r.n_blocks, "dst_off": r.dst_off, "out_len": r.out_len,
                        "message": lib().lz4ada_batch_message(self._h, k).decode()})
        return out

    def exact_sizing(self):
        """Size every block with K5 at upload time and place the blocks back to back (before upload())."""
        if lib().lz4ada_batch_exact_sizing(self._h) != 0:
            raise Assertion_Error("raised LZ4ADA.ASSERTION_ERROR : exact_sizing after upload")

    def set_output_capacity(self, nbytes):
        """Tell the batch how large the output buffer really is (room for streams that outgrow their region)."""
        lib().lz4ada_batch_set_output_capacity(self._h, nbytes)

    def retried_streams(self):
        return lib().lz4ada_batch_retried_streams(self._h)

    def k1_kernel_name(self):
        """The K1 kernel the last run() launched."""
        return lib().lz4ada_batch_k1_kernel_name(self._h).decode()

    def kernel_ms(self):
        """Device time of K1 / K4 / K3 in the last run (CUDA events on the launching stream)."""
        ms = (ctypes.c_float * 3)()
        lib().lz4ada_batch_kernel_ms(self._h, ms)
        return {"k1_decode_blocks": ms[0], "k4_decode_linked": ms[1], "k3_xxh32_frames": ms[2]}

    def traffic(self):
        a, b, c = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        lib().lz4ada_batch_traffic(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"compressed_read": a.value, "decompressed_written": b.value, "checksum_reread": c.value}

    def close(self):
        if self._h:
            lib().lz4ada_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def batch_decompress(ctx, streams, Reservation="For_All", exact_sizing=False, info=None):
    """Convenience over Batch for host data: list of bytes -> list of (exception, output, eof, message).
    exact_sizing: size every block first (K5) and place the blocks back to back.  info: optional dict that
    receives {"output_bytes", "retried_streams"} of the run."""
    offs, pos = [], 0
    for s in streams:
        offs.append((pos, len(s)))
        pos += len(s)
    src = b"".join(bytes(s) for s in streams) or b"\0"
    b = Batch(ctx, src, offs, Reservation)
    if exact_sizing:
        b.exact_sizing()
    need = b.output_bytes
    spare = 16 << 20   # room for crafted streams whose blocks inflate past the declared block maximum
    d_src = ctx.alloc(len(src) + 64)
    d_dst = ctx.alloc(need + spare + 64)
    b.set_output_capacity(need + spare)
    try:
        b.upload(d_src)
        b.run(d_src, d_dst)
        if info is not None:
            info["output_bytes"] = b.output_bytes
            info["retried_streams"] = b.retried_streams()
        out = []
        for r in b.results():
            data = ctx.d2h(d_dst + r["dst_off"], r["out_len"]) if r["out_len"] else b""
            out.append((r["exception"], data, r["end_of_frame"], r["message"]))
        return out
    finally:
        b.close()
        ctx.free(d_src)
        ctx.free(d_dst)
