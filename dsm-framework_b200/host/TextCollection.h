// TextCollection.h -- host-side mirror of the reference's index facade for the
// build-and-save half of the `builder` path (reference: TextCollection.h:37-129,
// FMIndex.h:171-222, FMIndex.cpp:155-217).  Objects are made by
// TextCollectionBuilder::InitTextCollection and own the finished index sections,
// which were produced on the GPU behind the C ABI of include/dsmfm.h.
//
// Same public names, argument meaning and error behaviour as the reference for
// everything `builder.cpp` uses; the query half (LF, getL, ...; FMIndex.h:61-154)
// is consumed from the saved `.fmi` by the reference's own metaenumerate and is
// not part of this path.
#ifndef DSMFM_HOST_TEXTCOLLECTION_H_
#define DSMFM_HOST_TEXTCOLLECTION_H_

#include <string>
#include <utility>
#include <vector>

#ifndef uchar
#define uchar unsigned char
#endif
#ifndef ulong
#define ulong unsigned long
#endif

struct dsmfm_builder;
struct dsmfm_index;

class TextCollection
{
public:
    typedef unsigned DocId;
    typedef ulong TextPosition;

    enum IndexType { TYPE_FMINDEX, TYPE_RLCSA };
    static const std::string REVERSE_EXTENSION;
    static const std::string ROTATION_EXTENSION;
    static const std::string FMINDEX_EXTENSION;
    static const std::string RLCSA_EXTENSION;

    // Total length of the indexed text including the 0-terminators.
    TextPosition getLength() const;
    DocId getNumberOfTexts() const;
    TextPosition getMaxTextLength() const;

    // Writes `<filename>.fmi` in the reference's version-17 layout.  Throws
    // std::runtime_error on an i/o error, like FMIndex::save.
    void save(std::string const &filename) const;

    // Writes `<filename>.sa` (FMIndex::saveSamples, FMIndex.cpp:125-147).  Available when the builder
    // kept the suffix array (environment DSMFM_KEEP_SA=1 / `builder --samples`); throws std::runtime_error otherwise.
    void saveSamples(std::string const &filename) const;

    ~TextCollection();

    bool isColorCoded() const { return colorCoded; }
    unsigned getRotationLength() const { return rotationLength; }

    // Device-side measurements of the build that produced this index (not in the reference).
    std::string buildReport() const;

private:
    friend class TextCollectionBuilder;
    TextCollection(dsmfm_builder *owner, dsmfm_index const &idx, bool cc, unsigned rl);
    TextCollection(TextCollection const &);
    TextCollection &operator=(TextCollection const &);

    dsmfm_builder *owner;   // keeps the host copies of the sections alive
    dsmfm_index *index;
    bool colorCoded;
    unsigned rotationLength;
};

#endif
