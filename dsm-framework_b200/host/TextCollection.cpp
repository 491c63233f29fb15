#include "TextCollection.h"
#include "dsmfm.h"

#include <cstdio>
#include <stdexcept>

const std::string TextCollection::REVERSE_EXTENSION = ".reverse";
const std::string TextCollection::ROTATION_EXTENSION = ".rotation";
const std::string TextCollection::FMINDEX_EXTENSION = ".fmi";
const std::string TextCollection::RLCSA_EXTENSION = ".rlcsa.array";

TextCollection::TextCollection(dsmfm_builder *owner_, dsmfm_index const &idx, bool cc, unsigned rl)
    : owner(owner_), index(new dsmfm_index(idx)), colorCoded(cc), rotationLength(rl)
{
}

TextCollection::~TextCollection()
{
    delete index;
    dsmfm_destroy(owner);
}

TextCollection::TextPosition TextCollection::getLength() const { return index->n; }
TextCollection::DocId TextCollection::getNumberOfTexts() const { return index->number_of_texts; }
TextCollection::TextPosition TextCollection::getMaxTextLength() const { return index->max_text_length; }

void TextCollection::save(std::string const &filename) const
{
    // colorCoded / rotationLength are always false / 0 on the builder path
    // (builder.cpp:19-22 never sets them); the C ABI writes them as such.
    if (dsmfm_write_fmi(index, filename.c_str()) != DSMFM_OK)
        throw std::runtime_error("TextCollection::save(): file write error.");
}

void TextCollection::saveSamples(std::string const &filename) const
{
    if (dsmfm_write_sa(owner, filename.c_str()) != DSMFM_OK)
        throw std::runtime_error(std::string("TextCollection::saveSamples(): ") + dsmfm_last_error(owner));
}

std::string TextCollection::buildReport() const
{
    dsmfm_stats s;
    if (dsmfm_get_stats(owner, &s) != DSMFM_OK) return "";
    char buf[512];
    std::snprintf(buf, sizeof buf,
                  "GPU build: n=%llu symbols, %u bits/symbol, %u sort passes, %u refinement rounds, "
                  "%.1f ms (pack %.1f, sort %.1f, refine %.1f, bwt %.1f, wavelet %.1f), %u kernel launches",
                  (unsigned long long)s.n, s.bits_per_symbol, s.sort_passes, s.rounds, s.ms_total, s.ms_pack,
                  s.ms_sort, s.ms_refine, s.ms_bwt, s.ms_wt, s.kernel_launches);
    return buf;
}
