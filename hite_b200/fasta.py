"""FASTA I/O byte-compatible with the reference's ``read_fasta`` / ``store_fasta``
(module/Util.py:1650-1670, 1983-1988): record name = first token (split on ' ' then '\\t'), sequence
upper-cased with line breaks removed, records with an empty name or sequence dropped; output is
``>name\\nSEQ\\n`` with the sequence on one line."""
import os


def read_fasta(fasta_path):
    contignames = []
    contigs = {}
    if os.path.exists(fasta_path):
        name, chunks = "", []
        with open(fasta_path, "r") as rf:
            for line in rf:
                if line.startswith(">"):
                    seq = "".join(chunks)
                    if name != "" and seq != "":
                        contigs[name] = seq
                        contignames.append(name)
                    name = line.strip()[1:].split(" ")[0].split("\t")[0]
                    chunks = []
                else:
                    chunks.append(line.strip().upper())
        seq = "".join(chunks)
        if name != "" and seq != "":
            contigs[name] = seq
            contignames.append(name)
    return contignames, contigs


def store_fasta(contigs, file_path):
    with open(file_path, "w") as f_save:
        for name in contigs.keys():
            f_save.write(">" + name + "\n" + contigs[name] + "\n")


def file_exist(resut_file):
    """module/Util.py:2831-2852 (same rules: non-empty FASTA with >=1 record; other files need a
    non-comment line; directories must be non-empty)."""
    if os.path.isfile(resut_file):
        if os.path.getsize(resut_file) > 0:
            if resut_file.endswith(".fa") or resut_file.endswith(".fasta"):
                _, contigs = read_fasta(resut_file)
                return len(contigs) > 0
            with open(resut_file, "r") as f_r:
                for line in f_r:
                    if not line.startswith("#") and line.strip():
                        return True
            return False
        return False
    if os.path.isdir(resut_file):
        return len(os.listdir(resut_file)) > 0
    return False
