"""Small helpers with the semantics of skdownscale/pointwise_models/utils.py:46-53."""

from __future__ import annotations

from typing import Any


def default_none_kwargs(kwargs: dict[str, Any] | None, copy: bool = False) -> dict[str, Any]:
    """``None`` means "no keyword arguments" (utils.py:46-53)."""
    if kwargs is None:
        return {}
    return dict(kwargs) if copy else kwargs
