"""ctypes view of the C-ABI declared in include/dcsb200.h (libdcsb200.so)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCSB200_LIB") or os.path.join(HERE, "libdcsb200.so")

OS93A, OS93B, OS94, OS95 = 0x9301, 0x9302, 0x9400, 0x9500
OK, E_EMPTY, E_TRUNCATED, E_BANDTYPE, E_SHORT, E_STOPPED = 0, -1, -2, -3, -4, -5
E_ARG, E_CUDA, E_NOMEM = -16, -17, -18


class StreamDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("nbytes", C.c_uint32), ("os_version", C.c_uint16),
                ("master_volume", C.c_uint8), ("mixing_level", C.c_uint8),
                ("tail_frames", C.c_uint16), ("reserved", C.c_uint16)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("frames", C.c_uint32), ("frames_decoded", C.c_uint32),
                ("stream_bytes", C.c_uint32), ("checksum", C.c_uint64)]


class RomInfo(C.Structure):
    _fields_ = [("os_version", C.c_uint16), ("hw_version", C.c_uint8), ("n_channels", C.c_uint8),
                ("version_number", C.c_uint16), ("n_tracks", C.c_uint16), ("catalog_offset", C.c_uint32),
                ("post_code", C.c_int32), ("signature", C.c_char * 128)]


class TrackInfo(C.Structure):
    _fields_ = [("address", C.c_uint32), ("channel", C.c_int32), ("type", C.c_int32), ("defer_code", C.c_uint16),
                ("looping", C.c_uint8), ("reserved", C.c_uint8), ("time", C.c_uint32)]


class Opcode(C.Structure):
    _fields_ = [("offset", C.c_int32), ("nesting_level", C.c_int32), ("loop_parent", C.c_int32), ("delay_count", C.c_uint16),
                ("opcode", C.c_uint8), ("n_operand_bytes", C.c_uint8), ("operand_bytes", C.c_uint8 * 8),
                ("desc", C.c_char * 64), ("hex_desc", C.c_char * 40)]


class PortWrite(C.Structure):
    _fields_ = [("frame", C.c_uint32), ("byte", C.c_uint8), ("pad", C.c_uint8 * 3)]


class Timeline(C.Structure):
    _fields_ = [("writes", C.POINTER(PortWrite)), ("n_writes", C.c_uint32), ("n_frames", C.c_uint32),
                ("master_volume", C.c_uint8), ("pad", C.c_uint8 * 3)]


class TimelineResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("frames", C.c_uint32), ("checksum", C.c_uint64),
                ("n_host_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class StreamInfo(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("n_bytes", C.c_int32), ("stream_type", C.c_int32), ("stream_subtype", C.c_int32),
                ("status", C.c_int32), ("header", C.c_uint8 * 16)]


SYMBOLS = {
    "dcsb_write_wav": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t]),
    "dcsb_write_dcs_file": (C.c_int, [C.c_char_p, C.c_uint16, C.c_void_p, C.c_size_t]),
    "dcsb_read_dcs_file": (C.c_longlong, [C.c_char_p, C.POINTER(C.c_uint16), C.c_void_p, C.c_size_t]),
    "dcsb_partition_streams": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "dcsb_player_stream_info": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(StreamInfo)]),
    "dcsb_rom_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "dcsb_rom_destroy": (None, [C.c_void_p]),
    "dcsb_rom_add": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "dcsb_rom_load_zip": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "dcsb_rom_check": (C.c_int, [C.c_void_p]),
    "dcsb_rom_get_info": (C.c_int, [C.c_void_p, C.POINTER(RomInfo)]),
    "dcsb_rom_track_info": (C.c_int, [C.c_void_p, C.c_uint16, C.POINTER(TrackInfo)]),
    "dcsb_rom_decompile_track": (C.c_size_t, [C.c_void_p, C.c_uint16, C.c_void_p, C.c_size_t]),
    "dcsb_rom_list_streams": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dcsb_rom_pointer": (C.c_void_p, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "dcsb_rom_last_error": (C.c_char_p, [C.c_void_p]),
    "dcsb_player_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dcsb_player_destroy": (None, [C.c_void_p]),
    "dcsb_player_set_master_volume": (None, [C.c_void_p, C.c_int]),
    "dcsb_player_write_data_port": (None, [C.c_void_p, C.c_uint8]),
    "dcsb_player_add_track_command": (None, [C.c_void_p, C.c_uint16]),
    "dcsb_player_load_audio_stream": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_int]),
    "dcsb_player_clear_tracks": (None, [C.c_void_p]),
    "dcsb_player_is_stream_playing": (C.c_int, [C.c_void_p, C.c_int]),
    "dcsb_player_render": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "dcsb_player_host_bytes": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dcsb_render_timelines": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Timeline), C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.POINTER(TimelineResult)]),
    "dcsb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "dcsb_destroy": (None, [C.c_void_p]),
    "dcsb_last_error": (C.c_char_p, [C.c_void_p]),
    "dcsb_version": (C.c_char_p, []),
    "dcsb_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "dcsb_set_pipeline": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "dcsb_decode_streams": (C.c_int, [C.c_void_p, C.POINTER(StreamDesc), C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(Result)]),
    "dcsb_batch_create": (C.c_int, [C.c_void_p, C.POINTER(StreamDesc), C.c_size_t, C.POINTER(C.c_void_p)]),
    "dcsb_batch_destroy": (None, [C.c_void_p]),
    "dcsb_batch_total_samples": (C.c_uint64, [C.c_void_p]),
    "dcsb_batch_total_frames": (C.c_uint64, [C.c_void_p]),
    "dcsb_batch_compressed_bytes": (C.c_uint64, [C.c_void_p]),
    "dcsb_batch_pcm_offset": (C.c_uint64, [C.c_void_p, C.c_size_t]),
    "dcsb_batch_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcsb_stream_extent": (C.c_size_t, [C.c_void_p, C.c_int]),
    "dcsb_encode_bound": (C.c_uint64, [C.c_uint64]),
    "dcsb_encode_streams": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "dcsb_player_set_lookahead": (C.c_int, [C.c_void_p, C.c_uint32]),
    "dcsb_rom_zip_files": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dcsb_batch_launches": (C.c_int, [C.c_void_p]),
    "dcsb_batch_launch_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dcsb_batch_results": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Result)]),
    "dcsb_batch_read_pcm": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "dcsb_batch_device_pcm": (C.c_void_p, [C.c_void_p]),
    "dcsb_batch_last_kernel_ms": (C.c_float, [C.c_void_p, C.c_int]),
    "dcsb_batch_read_scan": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dcsb_master_multiplier": (C.c_uint16, [C.c_int]),
    "dcsb_level_multiplier": (C.c_uint16, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dcsb_gain_stage": (C.c_int, [C.POINTER(C.c_uint16), C.c_uint, C.c_uint, C.c_uint16, C.POINTER(C.c_uint16)]),
}

_lib = None


def lib():
    """Load libdcsb200.so; raises (no fallback) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("dcsexplorer_b200: %s is missing -- run __graft_entry__.build() "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib
