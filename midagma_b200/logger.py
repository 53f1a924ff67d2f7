"""Run telemetry for the B200 path: rows fan out to pluggable sinks.

The fork's ``src/logger.py`` defines the *interface* callers program against -- a ``LogConfig`` record, a
``build_default_logger`` helper and a ``StructuredLogger`` with ``emit(event, metrics)`` / ``close()`` / ``load()`` and
the attributes ``run_dir``, ``jsonl_path``, ``csv_path`` (reference: src/logger.py:12-46, 49-76, 78-182).  That
interface is kept so a ``LogConfig`` built for the reference drives this package unchanged; the machinery behind it is
organised differently: a row is built once and handed to a list of small sink objects (memory, JSON-lines file, CSV
file, console, user callback), each of which can also be used on its own.  ``DagmaLinear`` produces the
``minimize.checkpoint`` rows at its convergence checkpoints.
"""
from __future__ import annotations

import csv
import json
import logging
import os
import uuid
from dataclasses import dataclass, field
from datetime import datetime
from typing import Any, Callable, Dict, Iterable, List, Optional

import numpy as np

Row = Dict[str, Any]


@dataclass
class LogConfig:
    """Same field names and defaults as the reference record (src/logger.py:12-46), so configs interchange."""
    enabled: bool = True
    print_to_console: bool = False
    level: int = logging.INFO
    log_every: int = 200
    outer_log_every: int = 1
    store_csv: bool = False
    store_jsonl: bool = True
    csv_path: Optional[str] = None
    jsonl_path: Optional[str] = None
    root_dir: str = "logs"
    run_dir: Optional[str] = None
    run_name: Optional[str] = None
    meta: Dict[str, Any] = field(default_factory=dict)
    callback: Optional[Callable[[Row], None]] = None
    keep_in_memory: bool = True
    include_cfg: bool = True


def build_default_logger(name: str = "score_structure_learning", level: int = logging.INFO, stream: bool = True,
                         logfile: Optional[str] = None) -> logging.Logger:
    """``logging.getLogger(name)`` that does not propagate and owns at most one stream and one file handler."""
    log = logging.getLogger(name)
    log.setLevel(level)
    log.propagate = False
    wanted = []
    if stream and not any(type(h) is logging.StreamHandler for h in log.handlers):
        wanted.append(logging.StreamHandler())
    if logfile and not any(isinstance(h, logging.FileHandler) and h.baseFilename == os.path.abspath(logfile)
                           for h in log.handlers):
        wanted.append(logging.FileHandler(logfile, encoding="utf-8"))
    layout = logging.Formatter("[%(asctime)s][%(levelname)s] %(message)s", datefmt="%H:%M:%S")
    for h in wanted:
        h.setFormatter(layout)
        h.setLevel(level)
        log.addHandler(h)
    return log


# ------------------------------------------------------------------------------------------------ sinks
class MemorySink:
    def __init__(self):
        self.rows: List[Row] = []

    def write(self, row: Row) -> None:
        self.rows.append(row)

    def close(self) -> None:
        pass


class JsonlSink:
    """One JSON object per line, flushed per row so a crashed run keeps what it logged."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "a", encoding="utf-8")

    def write(self, row: Row) -> None:
        self._f.write(json.dumps(row, ensure_ascii=False))
        self._f.write("\n")
        self._f.flush()

    def close(self) -> None:
        if not self._f.closed:
            self._f.close()


class CsvSink:
    """Columns are fixed by the first row written (the reference's behaviour)."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "a", newline="", encoding="utf-8")
        self._writer: Optional[csv.DictWriter] = None

    def write(self, row: Row) -> None:
        if self._writer is None:
            self._writer = csv.DictWriter(self._f, fieldnames=list(row))
            self._writer.writeheader()
        self._writer.writerow(row)
        self._f.flush()

    def close(self) -> None:
        if not self._f.closed:
            self._f.close()


class ConsoleSink:
    def __init__(self, logger: logging.Logger, level: int):
        self.logger, self.level = logger, level

    def write(self, row: Row) -> None:
        body = ", ".join(f"{k}={v:.4e}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items() if k != "event")
        self.logger.log(self.level, "%s | %s", row.get("event"), body)

    def close(self) -> None:
        pass


class CallbackSink:
    """A user hook must never take the optimisation down with it."""

    def __init__(self, fn: Callable[[Row], None], logger: logging.Logger):
        self.fn, self.logger = fn, logger

    def write(self, row: Row) -> None:
        try:
            self.fn(row)
        except Exception:                                     # noqa: BLE001
            self.logger.exception("telemetry callback raised; row dropped for that sink")

    def close(self) -> None:
        pass


def _fresh_run_dir(cfg: LogConfig) -> str:
    label = (cfg.run_name or "run").replace(" ", "_")
    return os.path.join(cfg.root_dir, f"{datetime.now():%Y%m%d-%H%M%S}_{label}_{uuid.uuid4().hex[:6]}")


def _read_rows(path: str) -> List[Row]:
    if path.endswith(".jsonl"):
        with open(path, "r", encoding="utf-8") as f:
            return [json.loads(line) for line in f if line.strip()]
    with open(path, "r", encoding="utf-8", newline="") as f:
        return list(csv.DictReader(f))


def rows_to_columns(rows: Iterable[Row]) -> Dict[str, np.ndarray]:
    """Column view of a row list (keys in first-seen order; missing cells are None)."""
    rows = list(rows)
    keys = list(dict.fromkeys(k for r in rows for k in r))
    cols = {}
    for k in keys:
        cells = [r.get(k) for r in rows]
        numeric = all(isinstance(c, (int, float, np.integer, np.floating)) and not isinstance(c, bool) for c in cells)
        cols[k] = np.array(cells) if numeric else np.array(cells, dtype=object)
    return cols


# ------------------------------------------------------------------------------------------------ the logger
class StructuredLogger:
    """``emit(event, metrics)`` -> one row ``{"event": event, **metrics}`` to every configured sink; nothing at all
    happens (no folder, no file, no row) while ``cfg.enabled`` is false."""

    def __init__(self, logger: logging.Logger, cfg: LogConfig):
        self.logger, self.cfg = logger, cfg
        self.run_dir: Optional[str] = None
        self.jsonl_path: Optional[str] = None
        self.csv_path: Optional[str] = None
        self._memory = MemorySink() if cfg.keep_in_memory else None
        self._sinks: list = []
        if not cfg.enabled:
            return
        if self._memory is not None:
            self._sinks.append(self._memory)
        if cfg.print_to_console:
            self._sinks.append(ConsoleSink(logger, cfg.level))
        if cfg.store_jsonl or cfg.store_csv:
            self.run_dir = cfg.run_dir or _fresh_run_dir(cfg)
            os.makedirs(self.run_dir, exist_ok=True)
            header = {"created_at": f"{datetime.now():%Y-%m-%d %H:%M:%S}", "run_name": cfg.run_name}
            header.update(cfg.meta or {})
            with open(os.path.join(self.run_dir, "meta.json"), "w", encoding="utf-8") as f:
                json.dump(header, f, ensure_ascii=False, indent=2)
            if cfg.store_jsonl:
                self.jsonl_path = cfg.jsonl_path or os.path.join(self.run_dir, "metrics.jsonl")
                self._sinks.append(JsonlSink(self.jsonl_path))
            if cfg.store_csv:
                self.csv_path = cfg.csv_path or os.path.join(self.run_dir, "metrics.csv")
                self._sinks.append(CsvSink(self.csv_path))
        if cfg.callback is not None:
            self._sinks.append(CallbackSink(cfg.callback, logger))

    @property
    def _rows(self) -> Optional[List[Row]]:
        return self._memory.rows if self._memory is not None else None

    def emit(self, event: str, metrics: Row) -> None:
        if not self._sinks:
            return
        row = {"event": event}
        row.update(metrics)
        for sink in self._sinks:
            sink.write(row)

    def close(self) -> None:
        for sink in self._sinks:
            sink.close()

    def load(self, *, source: Optional[str] = None, event: Optional[Any] = None) -> Dict[str, np.ndarray]:
        """Logged rows as column arrays -- from memory when there are any, else from ``source`` or the run's files."""
        if source is None and self._rows:
            rows = list(self._rows)
        else:
            path = source or self.jsonl_path or self.csv_path
            if path is None:
                raise ValueError("nothing to load: no rows in memory and no file sink")
            rows = _read_rows(path)
        if event is not None:
            keep = {event} if isinstance(event, str) else set(event)
            rows = [r for r in rows if r.get("event") in keep]
        return rows_to_columns(rows)
