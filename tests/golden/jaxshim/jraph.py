"""jraph stand-in (only imported, never used on the path).  Test infrastructure only."""
import collections

GraphsTuple = collections.namedtuple(
    "GraphsTuple", "nodes edges receivers senders globals n_node n_edge")
