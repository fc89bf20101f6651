{-# LANGUAGE DeriveGeneric #-}

-- |
-- Module      :  Data.BWT.Internal
-- Description :  drop-in replacement of text-compression's Data.BWT.Internal over the B200 kernels
--
-- Same export list, types and results as the reference module (src/Data/BWT/Internal.hs:42-56 of
-- Matthew-Mosior/text-compression v0.1.0.25); the suffix sort runs on the GPU (tc_bwt_encode) whenever the
-- input has at most 256 distinct elements, on ranks assigned in alphabet order
-- ("Data.TextCompression.Symbols").  The bodies are written against the semantics of the reference functions,
-- not taken from them.  NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.BWT.Internal ( -- * Base BWT types
                           Suffix(..),
                           SuffixArray,
                           BWT(..),
                           BWTMatrix(..),
                           -- * To BWT functions
                           saToBWT,
                           createSuffixArray,
                           -- * From BWT functions
                           sortTB,
                           STBWTCounter,
                           magicInverseBWT,
                           -- * Create BWT Matrix function
                           createBWTMatrix
                         ) where

import           Data.List                    (sortBy)
import           Data.Maybe                   (fromJust, isNothing)
import           Data.Sequence                (Seq (..))
import qualified Data.Sequence                as DS
import           Data.STRef                   (STRef)
import qualified Data.TextCompression.B200    as B200
import qualified Data.TextCompression.Symbols as Sym
import           GHC.Generics                 (Generic)

-- | One suffix: its rank (1-based), its 1-based start position, and the suffix itself
-- ('Nothing' for the empty suffix that follows the last symbol).
data Suffix a = Suffix { suffixindex    :: Int
                       , suffixstartpos :: Int
                       , suffix         :: Maybe (Seq a)
                       }
  deriving (Eq,Ord,Show,Read,Generic)

type SuffixArray a = Seq (Suffix a)

newtype BWT a = BWT (Seq (Maybe a))
  deriving (Eq,Ord,Show,Read,Generic)

newtype BWTMatrix a = BWTMatrix (Seq (Seq (Maybe a)))
  deriving (Eq,Ord,Show,Read,Generic)

-- | Kept for source compatibility (the reference exports the type of its loop counter).
type STBWTCounter s a = STRef s Int

-- | BWT column from a suffix array: the symbol before each suffix, 'Nothing' before the whole text.
saToBWT :: SuffixArray a -> Seq a -> Seq (Maybe a)
saToBWT sa t = fmap before sa
  where
    before s | suffixstartpos s == 1 = Nothing
             | otherwise             = Just (DS.index t (suffixstartpos s - 2))

-- | All n+1 suffixes (the empty one included, rank 1) in lexicographic order.
createSuffixArray :: Ord a => Seq a -> SuffixArray a
createSuffixArray xs = DS.mapWithIndex mk starts
  where
    n      = DS.length xs
    starts = case Sym.alphabetOf xs of
               Just al -> B200.createSuffixArrayW8 (Sym.encodeBS al xs)      -- GPU suffix sort on the ranks
               Nothing -> DS.fromList (sortBy bySuffix [1 .. n + 1])          -- > 256 distinct symbols: host
    bySuffix p q = compare (DS.drop (p - 1) xs) (DS.drop (q - 1) xs)
    mk i p = Suffix { suffixindex    = i + 1
                    , suffixstartpos = p
                    , suffix         = if p == n + 1 then Nothing else Just (DS.drop (p - 1) xs) }

-- | Order of the (symbol, position) pairs 'Data.BWT.fromBWT' sorts: by symbol, ties by position.
sortTB :: (Ord a1,Ord a2) => (a1, a2) -> (a1, a2) -> Ordering
sortTB (c1,i1) (c2,i2) = compare c1 c2 <> compare i1 i2

-- | Inverse BWT over the sorted (symbol, position) pairs: start at the pair that carries the position of the
-- 'Nothing' and follow the positions until the row of the 'Nothing' itself comes up again.
magicInverseBWT :: Seq (Maybe a,Int) -> Seq a
magicInverseBWT sorted =
  case DS.findIndexL (isNothing . fst) sorted of
    Nothing -> DS.empty
    Just e  -> go e (snd (DS.index sorted e)) DS.empty
  where
    go e f acc
      | f == e    = acc
      | otherwise = let (c, nxt) = DS.index sorted f in go e nxt (acc DS.|> fromJust c)

-- | Every rotation of @text ++ [Nothing]@, sorted: row k is the rotation that starts with the suffix of rank k.
-- Rows are built lazily from the (GPU) suffix array, so asking for the first column only -- what the FM-index
-- builders do -- does not materialise the O(n^2) matrix.
createBWTMatrix :: Ord a => [a] -> BWTMatrix a
createBWTMatrix [] = BWTMatrix DS.Empty
createBWTMatrix t  = BWTMatrix (fmap (row . suffixstartpos) (createSuffixArray s))
  where
    s     = DS.fromList t
    justs = fmap Just s
    row p = (DS.drop (p - 1) justs DS.|> Nothing) DS.>< DS.take (p - 1) justs
