"""Singleton logging wrapper with the reference's interface (logger/main_logger.py:9-103):
MainLogger(args).debug/info/warning/error/exception(msg, gpu_rank=-1).  The reference's `gpu_rank`
argument is a dead stub (its rank check always returns True); here it is live: in a multi-process
run only rank 0 emits, unless a message is addressed to a specific rank."""
import logging
import os
import sys
from datetime import datetime


class MainLogger:
    _instance = None
    _initialized = False

    def __new__(cls, *args, **kwargs):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    def __init__(self, args=None):
        if self._initialized:
            return
        self.logger_name = 'main'
        self.rank = int(os.environ.get('RANK', '0'))
        self.logger = logging.getLogger(self.logger_name)
        self.logger.setLevel(logging.DEBUG)
        fmt = logging.Formatter("%(asctime)s %(levelname)s:%(message)s")
        stream = logging.StreamHandler()
        stream.setFormatter(fmt)
        self.logger.addHandler(stream)
        if args is not None and getattr(args, 'log_file', 0) == 1 and self.rank == 0:
            path = args.save_path
            os.makedirs(path, exist_ok=True)
            fh = logging.FileHandler(os.path.join(path, f'{datetime.now().strftime("%Y%m%d_%H%M%S")}.log'))
            fh.setLevel(logging.DEBUG)
            fh.setFormatter(fmt)
            self.logger.addHandler(fh)
        self._initialized = True

        def catch_exception(exc_type, exc_value, exc_traceback):
            if issubclass(exc_type, KeyboardInterrupt):
                sys.__excepthook__(exc_type, exc_value, exc_traceback)
                return
            logging.getLogger("main").error("Unexpected exception.", exc_info=(exc_type, exc_value, exc_traceback))

        sys.excepthook = catch_exception

    def _emit(self, gpu_rank: int) -> bool:
        return self.rank == 0 if gpu_rank < 0 else self.rank == gpu_rank

    def debug(self, msg: str, gpu_rank: int = -1):
        if self._emit(gpu_rank):
            self.logger.debug(msg)

    def info(self, msg: str, gpu_rank: int = -1):
        if self._emit(gpu_rank):
            self.logger.info(msg)

    def warning(self, msg: str, gpu_rank: int = -1):
        if self._emit(gpu_rank):
            self.logger.warning(msg)

    def error(self, msg: str, gpu_rank: int = -1):
        if self._emit(gpu_rank):
            self.logger.error(msg)

    def exception(self, msg: str, gpu_rank: int = -1):
        if self._emit(gpu_rank):
            self.logger.exception(msg)
