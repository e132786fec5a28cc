"""`kmap` command line for the counting path: preproc / scan_motif / ex_hamball with the reference's option names
(reference cli.py:29-31, kmer_count.py:69-101, motif_discovery.py:21-52, 72-110).  The plotting / alignment /
visualisation commands of the reference read the files these commands write and stay in the reference package."""
import click


@click.group()
def cli():
    """KMAP counting path on B200 (k-mer counting, Hamming balls, distance matrix)."""


@cli.command(name="preproc")
@click.option("--fasta_file", type=str, required=True, help="Input fasta file")
@click.option("--res_dir", type=str, default=".", required=False, help="Result directory for storing all outputs")
@click.option("--gpu_mode", type=bool, default=False, required=False, help="accepted for compatibility (always GPU)")
@click.option("--debug", type=bool, default=False, required=False, help="display debug information.")
def preproc(fasta_file: str, res_dir=".", gpu_mode=False, debug=False):
    from .kmer_count import _preproc
    _preproc(fasta_file, res_dir, debug)


@cli.command(name="scan_motif")
@click.option("--res_dir", type=str, required=True, help="Result directory for storing all outputs")
@click.option("--gpu_mode", type=bool, default=False, required=False, help="accepted for compatibility (always GPU)")
@click.option("--debug", type=bool, default=False, required=False, help="display debug information.")
def scan_motif(res_dir: str, gpu_mode=False, debug=False):
    from .motif_discovery import _scan_motif
    _scan_motif(res_dir, debug)


@cli.command(name="ex_hamball")
@click.option("--res_dir", type=str, required=True, help="Result directory for storing all outputs")
@click.option("--conseq", type=str, required=True, help="the consensus sequence")
@click.option("--return_type", type=str, required=True, help='output file form, can be ["hash" | "kmer" | "matrix"]')
@click.option("--output_file", type=str, required=True, help="output file name, including the suffix")
@click.option("--max_ham_dist", type=int, default=-1, required=False,
              help="The radius of the Hamming ball. -1 means taking the radius from motif_def_table.csv")
def ex_hamball(res_dir: str, conseq: str, return_type: str, output_file: str, max_ham_dist: int = -1):
    from .motif_discovery import _ex_hamball
    _ex_hamball(res_dir, conseq, return_type, output_file, max_ham_dist)


def main():
    cli(prog_name="kmap")
