"""`validation` command placeholder.  The reference's validation (src/validation.py:29-99) scores
checkpoints with the training loss on random crops; it does not call predict().  Running it on the
fused kernel (batched, no prior, d=22) is row N1 of SURVEY.md section 8(f) -- next, not built yet."""
import click


@click.command(name='validation')
@click.option('--data', '-d', type=click.Path(file_okay=False, dir_okay=True), required=False)
@click.option('--checkpoints', '-c', type=click.Path(file_okay=False, dir_okay=True), required=False)
@click.option('--output', '-o', type=click.Path(), required=False)
def validation_command(data, checkpoints, output):
    raise click.ClickException('the `validation` command is not part of this build yet (SURVEY.md 8f, row N1); '
                               'the `inference` command is the supported hot path')
