/*
 * ref_shim.cpp -- thin extern "C" wrapper so that tests can call the UNMODIFIED reference
 * codebook builder.  It #includes huffTree.h from the reference checkout (-I$(REF)); no
 * reference source is copied into this repository.  TEST INFRASTRUCTURE ONLY.
 *
 * The call sequence mirrors loadData (load_data.h:34-47): BuildTree -> GenerateCodes ->
 * flatten each vector<bool> into a right-aligned codeword whose first tree edge is the MSB.
 * load_data.h itself cannot be included (it pulls in hist.cu and is missing a ';' at :28).
 */
#include <vector>
#include <cstring>
#include "huffTree.h"

extern "C" int ref_build_codebook(const unsigned int *freqs_in, unsigned int *codewords,
                                  unsigned int *codewordlens)
{
    unsigned int freqs[UniqueSymbols];
    int present = 0;
    for (int i = 0; i < UniqueSymbols; i++) { freqs[i] = freqs_in[i]; present += freqs[i] != 0; }
    std::memset(codewords, 0, 256 * sizeof(unsigned int));
    std::memset(codewordlens, 0, 256 * sizeof(unsigned int));
    if (present == 0) return 0;            /* reference would call top() on an empty queue */

    INode *root = BuildTree(freqs);
    HuffCodeMap codes;
    GenerateCodes(root, HuffCode(), codes);
    delete root;

    int maxlen = 0;
    for (HuffCodeMap::const_iterator it = codes.begin(); it != codes.end(); ++it) {
        unsigned int count = (unsigned int)it->second.size();
        if (count > 32) return -1;
        unsigned int cw = 0;
        for (unsigned int i = 0; i < count; i++)
            if (it->second[i]) cw += 1u << (count - i - 1);     /* == (uint)pow(2.0f, count-i-1) */
        codewords[(unsigned int)it->first] = cw;
        codewordlens[(unsigned int)it->first] = count;
        if ((int)count > maxlen) maxlen = (int)count;
    }
    return maxlen;
}
