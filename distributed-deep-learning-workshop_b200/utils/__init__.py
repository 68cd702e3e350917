"""Auxiliary subsystems: timeline tracer, clock sampler, bucket checksums."""
from .timeline import Timeline, timeline_from_env
from .debug import checksum_across_ranks
from .cpus import usable_cpus

__all__ = ["Timeline", "timeline_from_env", "checksum_across_ranks", "usable_cpus"]
