"""CLI entry point with the reference's command names (main.py:12-23).  Only the commands on or
next to the propagation hot path are registered: `inference`, `validation` and `evaluation` (`train` is out of scope)."""
import click

from src import evaluation, inference, validation

cli = click.Group(name='cli', commands={cmd.name: cmd for cmd in (inference.inference_command,
                                                                 validation.validation_command,
                                                                 evaluation.evaluation_command)})

if __name__ == '__main__':
    cli()
