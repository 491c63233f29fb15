#include "TextCollectionBuilder.h"
#include "dsmfm.h"

#include <cstdlib>
#include <iostream>
#include <vector>

struct TCBuilderRep
{
    TextCollection::IndexType type;
    unsigned samplerate;
    dsmfm_builder *gpu;
    ulong n;
    unsigned numberOfTexts;
    ulong maxTextLength;
    bool insertAllowed;
    bool lengthsCounted; // false once InsertFasta was used: the longest document is then only known to the device
    std::vector<std::string> name; // accepted and ignored by the index, as in the reference (FMIndex.cpp:100-116)
};

TextCollectionBuilder::TextCollectionBuilder(unsigned samplerate, ulong estimatedInputLength, TextCollection::IndexType type)
    : p_(new TCBuilderRep())
{
    p_->type = type;
    p_->samplerate = samplerate ? samplerate : TEXTCOLLECTION_DEFAULT_SAMPLERATE;
    p_->n = 0;
    p_->numberOfTexts = 0;
    p_->maxTextLength = 0;
    p_->insertAllowed = true;
    p_->lengthsCounted = true;
    p_->gpu = 0;

    dsmfm_options opt;
    std::memset(&opt, 0, sizeof opt);
    opt.device = -1;
    opt.samplerate = p_->samplerate;
    // builder.cpp:411 always passes an estimate of 1; anything below the default is "unknown"
    opt.expected_bytes = estimatedInputLength > TEXTCOLLECTION_DEFAULT_INPUT_LENGTH ? estimatedInputLength : 0;
    if (const char *dev = std::getenv("DSMFM_DEVICE")) opt.device = std::atoi(dev);
    if (const char *k = std::getenv("DSMFM_KEEP_SA"))
        if (std::atoi(k)) opt.flags |= DSMFM_FLAG_KEEP_SA; // TextCollection::saveSamples will be called
    if (dsmfm_create(&opt, &p_->gpu) != DSMFM_OK)
    {
        // no CPU fallback: without the GPU path there is no builder
        std::cerr << "TextCollectionBuilder: " << dsmfm_last_error(0) << std::endl;
        std::exit(1);
    }
}

TextCollectionBuilder::~TextCollectionBuilder()
{
    if (p_->gpu) dsmfm_destroy(p_->gpu); // not handed over to a TextCollection
    delete p_;
}

void TextCollectionBuilder::InsertText(uchar const *text)
{
    if (!p_->insertAllowed)
    {
        std::cerr << "TextCollectionBuilder::InsertText() error: new text can not be inserted after InitTextCollection() call!" << std::endl;
        std::exit(1);
    }
    ulong m = std::strlen((char const *)text) + 1;
    if (m > p_->maxTextLength) p_->maxTextLength = m;
    if (m <= 1)
    {
        std::cerr << "TextCollectionBuilder::InsertText() error: can not index empty texts!" << std::endl;
        std::exit(1);
    }
    p_->n += m;
    p_->numberOfTexts++;
    if (dsmfm_append(p_->gpu, text, m - 1) != DSMFM_OK)
    {
        std::cerr << "TextCollectionBuilder::InsertText() error: " << dsmfm_last_error(p_->gpu) << std::endl;
        std::exit(1);
    }
}

void TextCollectionBuilder::InsertText(uchar const *text, std::string const &name)
{
    p_->name.push_back(name);
    InsertText(text);
}

void TextCollectionBuilder::InsertFasta(uchar const *text, ulong length, bool final, FastaReport &report)
{
    if (!p_->insertAllowed)
    {
        std::cerr << "TextCollectionBuilder::InsertFasta() error: new text can not be inserted after InitTextCollection() call!" << std::endl;
        std::exit(1);
    }
    dsmfm_fasta_info info;
    if (dsmfm_append_fasta(p_->gpu, text, length, final ? 1 : 0, &info) != DSMFM_OK)
    {
        std::cerr << "TextCollectionBuilder::InsertFasta() error: " << dsmfm_last_error(p_->gpu) << std::endl;
        std::exit(1);
    }
    p_->n += info.doc_bytes;
    p_->numberOfTexts += (unsigned)info.documents;
    p_->lengthsCounted = false;
    report.consumed = info.consumed;
    report.records = info.records;
    report.documents = info.documents;
    report.bases = info.bases;
    report.invalidRecords = info.invalid_records;
    report.firstInvalidOffset = info.first_invalid_offset;
    report.badHeaders = info.bad_headers;
}

void *TextCollectionBuilder::AllocPinned(ulong bytes) { return dsmfm_alloc_pinned(bytes); }
void TextCollectionBuilder::FreePinned(void *p) { dsmfm_free_pinned(p); }

TextCollection *TextCollectionBuilder::InitTextCollection(bool storePlainText, bool color, unsigned rotationLength)
{
    p_->insertAllowed = false;
    (void)storePlainText; // the reference never stores plain text on this path (builder.cpp:273-276)
    switch (p_->type)
    {
    case TextCollection::TYPE_FMINDEX:
    {
        dsmfm_index idx;
        if (dsmfm_finish(p_->gpu, &idx) != DSMFM_OK)
        {
            std::cerr << "TextCollectionBuilder::InitTextCollection() error: " << dsmfm_last_error(p_->gpu) << std::endl;
            std::exit(1);
        }
        // what the device found must be what InsertText counted (the reference asserts length == n,
        // TextCollectionBuilder.cpp:128)
        if (p_->numberOfTexts != 0 &&
            (idx.n != p_->n || idx.number_of_texts != p_->numberOfTexts ||
             (p_->lengthsCounted && idx.max_text_length != p_->maxTextLength)))
        {
            std::cerr << "TextCollectionBuilder::InitTextCollection() error: device/host document accounting mismatch" << std::endl;
            std::exit(1);
        }
        TextCollection *result = new TextCollection(p_->gpu, idx, color, rotationLength);
        p_->gpu = 0; // owned by the collection now
        return result;
    }
    case TextCollection::TYPE_RLCSA:
        std::cerr << "TextCollectionBuilder::InitTextCollection(): currently unsupported!" << std::endl;
        std::abort();
    default:
        std::cerr << "TextCollectionBuilder::InitTextCollection(): invalid index type!" << std::endl;
        std::exit(2);
    }
    return 0;
}
