{-# LANGUAGE DeriveGeneric #-}

-- |
-- Module      :  Data.FMIndex.Internal
-- Description :  drop-in replacement of text-compression's Data.FMIndex.Internal
--
-- Export list and types of the reference (src/Data/FMIndex/Internal.hs:120-135).  The 'FMIndex' value keeps the
-- reference's shape -- C table, dense Occ(c,k) table, full suffix array -- so code that pattern matches on it
-- keeps working; it is only practical for small inputs (sigma x N boxed triples), exactly as in the reference.
-- The batch functions of "Data.FMIndex" do not go through this value: they search the device-resident index
-- (rank blocks + sampled suffix array).  'countFMIndex' / 'locateFMIndex' below are the backward search over
-- the dense value, including the reference's handling of pattern symbols that do not occur in the text
-- (the search stops and keeps the range found so far: SURVEY.md 2.3 Q4).
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.FMIndex.Internal ( -- * Base FM-index types
                               FMIndex(..),
                               OccCK(..),
                               Cc(..),
                               SA(..),
                               -- * To OccCK (ByteString) functions
                               seqToOccCK,
                               -- * Cc (ByteString) functions
                               seqToCc,
                               -- * From FMIndex (ByteString) functions
                               seqFromFMIndex,
                               -- * Count (ByteString) operation
                               countFMIndex,
                               -- * Locate (ByteString) operation
                               locateFMIndex,
                             ) where

import           Data.BWT.Internal (SuffixArray)
import           Data.MTF.Internal (nubSeq')
import           Data.RLE.Internal (Pack)

import           Data.Foldable     (toList)
import           Data.Sequence     (Seq (..))
import qualified Data.Sequence     as DS
import           GHC.Generics      (Generic)

newtype FMIndex b = FMIndex (Cc b,OccCK b,SA b)
  deriving (Eq,Ord,Show,Read,Generic)

-- | Per symbol of the alphabet: (k, Occ(c,k), BWT[k]) for k = 1..N, Occ inclusive of k.
newtype OccCK b = OccCK (Seq (Maybe b,Seq (Int,Int,Maybe b)))
  deriving (Eq,Ord,Show,Read,Generic)

-- | C[c]: number of symbols of the text$ smaller than c, per symbol of the sorted alphabet.
newtype Cc b = Cc (Seq (Int,Maybe b))
  deriving (Eq,Ord,Show,Read,Generic)

newtype SA b = SA (SuffixArray b)
  deriving (Eq,Ord,Show,Read,Generic)

-- | Dense Occ table of a BWT column.
seqToOccCK :: (Pack b, Ord b) => Seq (Maybe b) -> Seq (Maybe b,Seq (Int,Int,Maybe b))
seqToOccCK DS.Empty = DS.empty
seqToOccCK col      = fmap row (nubSeq' col)
  where
    row c = (c, DS.fromList (zip3 [1 ..] (drop 1 (scanl (\acc y -> if y == c then acc + 1 else acc) 0 ys)) ys))
    ys    = toList col

-- | C table of an F column (the sorted symbols): index of the first occurrence of every symbol.
seqToCc :: Ord b => Seq (Maybe b) -> Seq (Int,Maybe b)
seqToCc DS.Empty = DS.empty
seqToCc f        = fmap (\c -> (maybe 0 id (DS.elemIndexL c f), c)) (nubSeq' f)

-- | The BWT column stored in the index (third component of any Occ row).
seqFromFMIndex :: FMIndex b -> Seq (Maybe b)
seqFromFMIndex (FMIndex (Cc DS.Empty,_,_))    = DS.Empty
seqFromFMIndex (FMIndex (_,OccCK DS.Empty,_)) = DS.Empty
seqFromFMIndex (FMIndex (_,OccCK (r :<| _),_)) = fmap (\(_,_,s) -> s) (snd r)

-- | 1-based inclusive SA range of a pattern, or 'Nothing'.
searchFMIndex :: Eq b => Seq b -> FMIndex b -> Maybe (Int, Int)
searchFMIndex DS.Empty _ = Nothing
searchFMIndex _ (FMIndex (Cc DS.Empty,_,_)) = Nothing
searchFMIndex pat (FMIndex (Cc cc,OccCK occ,_)) = go (reverse (toList pat)) Nothing
  where
    n          = maybe 0 (DS.length . snd) (DS.lookup 0 occ)
    cOf a      = DS.findIndexL ((== Just a) . snd) cc
    occAt j k  = if k <= 0 then 0 else let (_,o,_) = DS.index (snd (DS.index occ j)) (k - 1) in o
    finish r   = case r of
                   Just (s, e) | e >= s -> Just (s, e)
                   _                    -> Nothing
    go [] r = finish r
    go (a : as) r
      | Just (s, e) <- r, s > e = Nothing
      | otherwise = case cOf a of
          Nothing -> finish r                                  -- symbol not in the text: keep what was found
          Just j  ->
            let cj = fst (DS.index cc j)
            in case r of
                 Nothing     -> go as (Just (cj + 1, maybe n fst (DS.lookup (j + 1) cc)))
                 Just (s, e) -> go as (Just (cj + occAt j (s - 1) + 1, cj + occAt j e))

countFMIndex :: Pack b => Seq b -> FMIndex b -> Maybe Int
countFMIndex pat fm = fmap (\(s, e) -> e - s + 1) (searchFMIndex pat fm)

-- | The SA ranks (1-based) of the occurrences; the wrappers of "Data.FMIndex" map them to text positions.
locateFMIndex :: Eq b => Seq b -> FMIndex b -> Seq (Maybe Int)
locateFMIndex pat fm = maybe DS.empty (\(s, e) -> DS.fromList (map Just [s .. e])) (searchFMIndex pat fm)
