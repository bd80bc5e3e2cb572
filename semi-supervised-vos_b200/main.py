"""CLI entry point with the reference's command names (main.py:12-23).  Only the commands on or
next to the propagation hot path are registered: `inference`, `validation` and `evaluation` (`train` is out of scope)."""
import click

from src.evaluation import evaluation_command
from src.inference import inference_command
from src.validation import validation_command


@click.group(name='cli')
def cli():
    pass


cli.add_command(inference_command)
cli.add_command(validation_command)
cli.add_command(evaluation_command)

if __name__ == '__main__':
    cli()
