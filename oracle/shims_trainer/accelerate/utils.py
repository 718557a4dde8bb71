"""ORACLE / TEST INFRASTRUCTURE ONLY."""


def broadcast(t, from_process=0):
    return t
