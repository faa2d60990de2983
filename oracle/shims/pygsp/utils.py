import logging


def build_logger(name):
    return logging.getLogger(name)
