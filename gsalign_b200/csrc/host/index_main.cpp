// bin/gsa_index ref.fa prefix -- drop-in for the reference's bin/bwt_index (src/BWT_Index/main.c:5-14),
// built on the GPU suffix sorter in index_build.cu.  Writes prefix.{pac,ann,amb,bwt,sa} in the BWA format.
#include <stdio.h>
extern "C" int gsa_build_index_files(const char *fasta, const char *prefix, int device);
int main(int argc, char *argv[])
{
	if (argc != 3) { fprintf(stderr, "usage: %s <Fasta_File> <Prefix>\n", argv[0]); return 0; }
	return gsa_build_index_files(argv[1], argv[2], 0) == 0 ? 0 : 1;
}
