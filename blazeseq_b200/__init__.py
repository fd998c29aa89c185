"""blazeseq_b200 -- B200-native FASTQ record tokenizer / validator / SoA packer.

A drop-in for the hot path of MoSafi2/BlazeSeq (FastqParser.views()/records()/batches()): the
host API mirrors the reference, the byte work runs in hand-written sm_100a CUDA kernels behind
the C ABI of include/blazeseq_gpu.h (blazeseq_b200/lib/libblazeseq_gpu.so).  There is no CPU
parsing path: without the shared library or a CUDA device, creating a parser raises.
"""
from . import _capi
from .host import (DEFAULT_BATCH_SIZE, DEFAULT_CAPACITY, EOF, MAX_CAPACITY, BlazeSeqError,
                     DeviceFastqBatch, EOFError, FastqBatch, FastqGZParser, FastqParser, FastqRecord,
                     FastqView, FileReader, GpuParser, GZFile, MemoryReader, ParserConfig,
                     QualitySchema, RapidgzipReader, Reader, create_parser, parse_schema, parser, shard_prefix)
from .fasta import FastaParser, FastaParserConfig, FastaRecord
from .pipeline import HostBatchPipeline

__all__ = [
    "DEFAULT_BATCH_SIZE", "DEFAULT_CAPACITY", "EOF", "MAX_CAPACITY", "BlazeSeqError",
    "DeviceFastqBatch", "EOFError", "FastqBatch", "FastqGZParser", "FastqParser", "FastqRecord",
    "FastqView", "FileReader", "GpuParser", "GZFile", "MemoryReader", "ParserConfig",
    "QualitySchema", "RapidgzipReader", "Reader", "create_parser", "parse_schema", "parser", "shard_prefix",
    "FastaParser", "FastaParserConfig", "FastaRecord", "HostBatchPipeline",
]
__version__ = "0.1.0"
