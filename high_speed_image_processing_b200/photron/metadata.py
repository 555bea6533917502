"""Which header fields a PhotonVideo exposes (reference: src/photron/metadata.py:11-129).

A pure allow-list over the decode seam's info dict; no performance relevance."""
from __future__ import annotations

from typing import FrozenSet, Iterable, Optional, Set


class MetadataConfig:
    ESSENTIAL: FrozenSet[str] = frozenset(
        {"Total Frame", "Image Width", "Image Height", "EffectiveBit Depth", "File Format"})
    RECORDING: FrozenSet[str] = frozenset({"Record Rate(fps)", "Shutter Speed(s)"})
    DEVICE: FrozenSet[str] = frozenset({"Camera Type", "Date"})
    EXTENDED: FrozenSet[str] = frozenset(
        {"Original Total Frame", "EffectiveBit Side", "Color Bit", "Comment Text"})
    ALL_FIELDS: FrozenSet[str] = ESSENTIAL | RECORDING | DEVICE | EXTENDED

    def __init__(self, fields: Optional[Iterable[str]] = None, include_essential: bool = True):
        chosen: Set[str] = set(self.ESSENTIAL) if include_essential else set()
        if fields is not None:
            chosen |= set(fields)
        self._fields = chosen

    @classmethod
    def minimal(cls) -> "MetadataConfig":
        return cls()

    @classmethod
    def full(cls) -> "MetadataConfig":
        return cls(fields=cls.ALL_FIELDS)

    @classmethod
    def for_processing(cls) -> "MetadataConfig":
        return cls(fields=cls.ESSENTIAL | cls.RECORDING)

    @property
    def fields(self) -> Set[str]:
        return set(self._fields)

    def should_include(self, field_name: str) -> bool:
        return field_name in self._fields

    def filter_metadata(self, raw_metadata: dict) -> dict:
        return {k: v for k, v in raw_metadata.items() if k in self._fields}

    def __repr__(self) -> str:
        return f"MetadataConfig(fields={sorted(self._fields)})"
