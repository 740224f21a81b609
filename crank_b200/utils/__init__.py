"""Mirror of the crank.utils functions the eval path needs on the device."""
from .griffin_lim import griffin_lim, logmelspc_to_linearspc, mlfb2wav  # noqa: F401
