"""Structured run telemetry with the interface of the fork's ``src/logger.py`` (reference: LogConfig :12-46,
build_default_logger :49-76, StructuredLogger :78-172): every ``emit(event, metrics)`` becomes one row
``{"event": ..., **metrics}`` that is kept in memory, appended to ``metrics.jsonl`` / ``metrics.csv`` in a per-run
folder, printed through a ``logging.Logger`` and / or handed to a callback -- each sink switched by ``LogConfig``.
The rows of ``minimize.checkpoint`` are produced by ``DagmaLinear`` at the convergence checkpoints.
"""
from __future__ import annotations

import csv
import json
import logging
import os
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import numpy as np


@dataclass
class LogConfig:
    """Field names and defaults of the reference's LogConfig (logger.py:12-46)."""
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
    callback: Optional[Callable[[Dict[str, Any]], None]] = None
    keep_in_memory: bool = True
    include_cfg: bool = True


def build_default_logger(name: str = "score_structure_learning", level: int = logging.INFO, stream: bool = True,
                         logfile: Optional[str] = None) -> logging.Logger:
    """A non-propagating logger with a stream (and optionally a file) handler, configured once per name."""
    log = logging.getLogger(name)
    log.setLevel(level)
    log.propagate = False
    if not getattr(log, "_configured", False):
        fmt = logging.Formatter("[%(asctime)s][%(levelname)s] %(message)s", datefmt="%H:%M:%S")
        handlers: List[logging.Handler] = []
        if stream:
            handlers.append(logging.StreamHandler())
        if logfile:
            handlers.append(logging.FileHandler(logfile, encoding="utf-8"))
        for h in handlers:
            h.setLevel(level)
            h.setFormatter(fmt)
            log.addHandler(h)
        log._configured = True
    return log


class StructuredLogger:
    """Row sink; ``emit`` is a no-op when ``cfg.enabled`` is false (logger.py:140-172)."""

    def __init__(self, logger: logging.Logger, cfg: LogConfig):
        self.logger, self.cfg = logger, cfg
        self._rows: Optional[List[Dict[str, Any]]] = [] if cfg.keep_in_memory else None
        self.run_dir = self.jsonl_path = self.csv_path = None
        self._jsonl_f = self._csv_f = None
        self._csv_header_written = False
        to_disk = cfg.enabled and (cfg.store_csv or cfg.store_jsonl)
        if to_disk:
            self.run_dir = cfg.run_dir or self._new_run_dir()
            os.makedirs(self.run_dir, exist_ok=True)
            with open(os.path.join(self.run_dir, "meta.json"), "w", encoding="utf-8") as f:
                json.dump({"created_at": time.strftime("%Y-%m-%d %H:%M:%S"), "run_name": cfg.run_name, **(cfg.meta or {})},
                          f, ensure_ascii=False, indent=2)
            if cfg.store_jsonl:
                self.jsonl_path = cfg.jsonl_path or os.path.join(self.run_dir, "metrics.jsonl")
                self._jsonl_f = open(self.jsonl_path, "a", encoding="utf-8")
            if cfg.store_csv:
                self.csv_path = cfg.csv_path or os.path.join(self.run_dir, "metrics.csv")
                self._csv_f = open(self.csv_path, "a", newline="", encoding="utf-8")

    def _new_run_dir(self) -> str:
        stamp = time.strftime("%Y%m%d-%H%M%S")
        name = (self.cfg.run_name or "run").replace(" ", "_")
        return os.path.join(self.cfg.root_dir, f"{stamp}_{name}_{int(time.time() * 1000) % 100000}")

    def close(self) -> None:
        for f in (self._jsonl_f, self._csv_f):
            if f:
                f.close()
        self._jsonl_f = self._csv_f = None

    def emit(self, event: str, metrics: Dict[str, Any]) -> None:
        if not self.cfg.enabled:
            return
        row = {"event": event, **metrics}
        if self._rows is not None:
            self._rows.append(row)
        if self.cfg.print_to_console:
            self.logger.log(self.cfg.level, f"{event} | " + self._fmt(metrics))
        if self._jsonl_f:
            self._jsonl_f.write(json.dumps(row, ensure_ascii=False) + "\n")
            self._jsonl_f.flush()
        if self._csv_f:
            w = csv.DictWriter(self._csv_f, fieldnames=list(row.keys()))
            if not self._csv_header_written:
                w.writeheader()
                self._csv_header_written = True
            w.writerow(row)
            self._csv_f.flush()
        if self.cfg.callback:
            try:
                self.cfg.callback(row)
            except Exception:
                self.logger.exception("logging callback failed")

    @staticmethod
    def _fmt(d: Dict[str, Any]) -> str:
        return ", ".join(f"{k}={v:.4e}" if isinstance(v, float) else f"{k}={v}" for k, v in d.items())

    def load(self, *, source: Optional[str] = None, event: Optional[Any] = None) -> Dict[str, np.ndarray]:
        """Rows as column arrays: the memory buffer first, else ``source`` / the jsonl / the csv file."""
        if source is None and self._rows:
            rows = list(self._rows)
        else:
            path = source or self.jsonl_path or self.csv_path
            if path is None:
                raise ValueError("nothing to load: no rows in memory and no file sink")
            if path.endswith(".jsonl"):
                with open(path, "r", encoding="utf-8") as f:
                    rows = [json.loads(line) for line in f if line.strip()]
            else:
                with open(path, "r", encoding="utf-8", newline="") as f:
                    rows = list(csv.DictReader(f))
        if event is not None:
            wanted = {event} if isinstance(event, str) else set(event)
            rows = [r for r in rows if r.get("event") in wanted]
        keys: List[str] = []
        for r in rows:
            for k in r:
                if k not in keys:
                    keys.append(k)
        return {k: np.array([r.get(k) for r in rows], dtype=object if any(isinstance(r.get(k), (dict, str, type(None)))
                                                                           for r in rows) else None) for k in keys}
